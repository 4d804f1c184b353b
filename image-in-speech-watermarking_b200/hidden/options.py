"""`HiDDenConfiguration` of the reference (`hidden/options.py:20-49`): plain configuration holder."""


class HiDDenConfiguration():
    def __init__(self, H: int, W: int, message_length: int, encoder_blocks: int, encoder_channels: int,
                 decoder_blocks: int, decoder_channels: int, use_discriminator: bool, use_vgg: bool,
                 discriminator_blocks: int, discriminator_channels: int, decoder_loss: float, encoder_loss: float,
                 adversarial_loss: float, enable_fp16: bool = False):
        self.H, self.W, self.message_length = H, W, message_length
        self.encoder_blocks, self.encoder_channels = encoder_blocks, encoder_channels
        self.use_discriminator, self.use_vgg = use_discriminator, use_vgg
        self.decoder_blocks, self.decoder_channels = decoder_blocks, decoder_channels
        self.discriminator_blocks, self.discriminator_channels = discriminator_blocks, discriminator_channels
        self.decoder_loss, self.encoder_loss, self.adversarial_loss = decoder_loss, encoder_loss, adversarial_loss
        self.enable_fp16 = enable_fp16
