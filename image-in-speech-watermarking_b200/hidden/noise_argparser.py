"""`--noise` mini-grammar of the reference (`hidden/noise_argparser.py:22-107`):

    'crop((0.4,0.55),(0.4,0.55))+cropout((0.25,0.35),(0.25,0.35))+dropout(0.25,0.35)+resize(0.4,0.6)+jpeg()+quant()'

-> a list of noise-layer modules / placeholders that `Noiser` consumes.  Same commands, same
placeholders ('JpegPlaceholder', 'QuantizationPlaceholder'), same ValueError on an unknown command."""
import argparse
import re

from .noise_layers import Crop, Cropout, Dropout, Resize

_NUM = r'(\d+\.*\d*),(\d+\.*\d*)'
_PAIR = re.compile(r'\w+\(\(' + _NUM + r'\),\(' + _NUM + r'\)\)')
_ONE = re.compile(r'\w+\(' + _NUM + r'\)')


def _two_ranges(command):
    m = _PAIR.match(command)
    if m is None:
        raise ValueError('Command not recognized: \n{}'.format(command))
    a, b, c, d = (float(v) for v in m.groups())
    return (a, b), (c, d)


def _one_range(command):
    m = _ONE.match(command)
    if m is None:
        raise ValueError('Command not recognized: \n{}'.format(command))
    a, b = (float(v) for v in m.groups())
    return (a, b)


def parse_noise(spec):
    """The list `NoiseArgParser.__call__` stores (`noise_argparser.py:81-107`)."""
    layers = []
    for command in spec.split('+'):
        command = command.replace(' ', '')
        if command.startswith('cropout'):
            layers.append(Cropout(*_two_ranges(command)))
        elif command.startswith('crop'):
            layers.append(Crop(*_two_ranges(command)))
        elif command.startswith('dropout'):
            layers.append(Dropout(_one_range(command)))
        elif command.startswith('resize'):
            layers.append(Resize(_one_range(command)))
        elif command.startswith('jpeg'):
            layers.append('JpegPlaceholder')
        elif command.startswith('quant'):
            layers.append('QuantizationPlaceholder')
        elif command.startswith('identity'):
            pass                                    # Noiser always holds one Identity()
        else:
            raise ValueError('Command not recognized: \n{}'.format(command))
    return layers


class NoiseArgParser(argparse.Action):
    def __call__(self, parser, namespace, values, option_string=None):
        setattr(namespace, self.dest, parse_noise(values[0]))
