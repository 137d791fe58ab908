"""``fma(a, b, c) = a * b + c`` (stylegan2_ada_pytorch/torch_utils/ops/fma.py:19-52): demodulation x activations + noise.
Elementwise glue on broadcasting tensors; PyTorch's own autograd already gives the cheap gradients the reference
hand-writes, so this is a plain expression (the heavy lifting of the layer is the convolution)."""


def fma(a, b, c):
    return a * b + c
