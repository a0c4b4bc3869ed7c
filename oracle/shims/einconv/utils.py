from typing import Tuple, Union


def get_conv_paddings(kernel_size: int, stride: int, padding: Union[int, str], dilation: int) -> Tuple[int, int]:
    """Left/right padding of a 1-d convolution for 'valid' / 'same' string paddings."""
    if isinstance(padding, int):
        return padding, padding
    if padding == "valid":
        return 0, 0
    if padding == "same":
        if stride != 1:
            raise ValueError("padding='same' requires stride 1")
        total = dilation * (kernel_size - 1)
        left = total // 2
        return left, total - left
    raise ValueError(f"Unknown padding {padding!r}")
