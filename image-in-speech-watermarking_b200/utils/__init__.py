from .model_utils import *  # noqa: F401,F403  (mirrors uformerWM/utils/__init__.py)
