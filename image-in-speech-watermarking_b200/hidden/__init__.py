"""HiDDeN flavour of the hot path (reference `hidden/`): the modified Decoder / ConvBNRelu and the
noise layers, executing on libwmk kernels."""
