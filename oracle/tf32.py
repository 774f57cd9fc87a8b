"""ORACLE (test infrastructure, never the product path): the arithmetic of `tcgen05.mma.kind::tf32` restated on
the CPU, so that the tensor-core kernels (csrc/neumf_tc.cu, csrc/neumf_fused.cu, csrc/gemm_tc.cu) are tested against
an oracle and not against the repository's own fp32 kernels.

The tensor core reads fp32 bit patterns from shared memory and uses sign, 8 exponent bits and the TOP 10 mantissa
bits of each operand; products are exact and are accumulated in fp32.  Whether the dropped 13 bits are truncated or
rounded is not written in the guides this repo has; both forms are here, `tests/test_gpu_tc.py` determines on the
device which one the hardware implements (truncation) and the parity tests use that one.

`Linear` is the autograd form of one Dense product as the kernels compute all three of its products:
    forward      y  = tf32(x) @ tf32(W)
    input grad   dx = tf32(dy) @ tf32(W)^T
    weight grad  dW = tf32(x)^T @ tf32(dy)
(every operand is rounded where it is READ by a product; sums stay fp32).  Passing `matmul=tf32.matmul` to
oracle/neumf.py or oracle/twotower.py turns those oracles into the TF32-operand oracle of the same graph.
"""
import numpy as np
import torch


def trunc_np(x):
    """fp32 -> TF32 by truncation: the low 13 mantissa bits cleared."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rn_np(x):
    """fp32 -> TF32 by round-to-nearest-even on the 13 dropped bits (what cvt.rna.tf32 would give, ties aside)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = x.view(np.uint32).astype(np.uint64)
    lsb = (b >> np.uint64(13)) & np.uint64(1)
    b = (b + np.uint64(0x0FFF) + lsb) & np.uint64(0xFFFFE000)
    return b.astype(np.uint32).view(np.float32)


def _round_t(x, mode):
    if x.dtype != torch.float32:
        return x                                    # the fp64 run of an oracle is the exact reference: no rounding
    a = x.detach().contiguous().view(torch.int32)
    if mode == "trunc":
        a = a & -8192                               # 0xFFFFE000
    else:
        lsb = (a >> 13) & 1
        a = (a + 0x0FFF + lsb) & -8192
    return a.view(torch.float32)


class Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, mode):
        ctx.mode = mode
        ctx.save_for_backward(x, w)
        return _round_t(x, mode) @ _round_t(w, mode)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        g = _round_t(gy, ctx.mode)
        return g @ _round_t(w, ctx.mode).t(), _round_t(x, ctx.mode).t() @ g, None


def matmul(x, w, mode="trunc"):
    return Linear.apply(x, w, mode)


def matmul_rn(x, w):
    return Linear.apply(x, w, "rn")
