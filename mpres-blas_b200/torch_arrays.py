"""mp_array_t views over torch-owned device memory + on-device synthetic inputs.

PyTorch is plumbing here: it owns the device allocations (so torch.distributed / NCCL can move the SoA
arrays of an mp_array_t between GPUs) and provides the random bits; the conversion to the RNS format is
this library's own kernel (mpres_array_set_binary).
"""
import ctypes

import torch

from . import MpresError, _check, mp_array_t


class TorchMpArray:
    """An mp_array_t (src/types.cuh:85-92) whose five device arrays are torch tensors."""

    def __init__(self, ctx, size, device=None):
        dev = torch.device("cuda", ctx.device) if device is None else device
        self.ctx, self.size = ctx, int(size)
        n = max(1, self.size)
        self.digits = torch.zeros(n * ctx.N, dtype=torch.int32, device=dev)
        self.sign = torch.zeros(n, dtype=torch.int32, device=dev)
        self.exp = torch.zeros(n, dtype=torch.int32, device=dev)
        self.eval = torch.zeros(2 * n * 2, dtype=torch.int64, device=dev)   # 2*len x {double frac; long exp}
        self.len = torch.tensor([self.size], dtype=torch.int32, device=dev)
        self.s = mp_array_t(self.digits.data_ptr(), self.sign.data_ptr(), self.exp.data_ptr(), self.eval.data_ptr(),
                            None, self.len.data_ptr())

    def tensors(self):
        """the SoA arrays that make up the number data (what a broadcast has to move)"""
        return [self.digits, self.sign, self.exp, self.eval]

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors())

    def host2device(self, recs):
        import numpy as np
        recs = np.ascontiguousarray(recs, dtype=self.ctx.dtype).reshape(-1)
        _check(self.ctx.lib.mpres_array_host2device(self.ctx.h, ctypes.byref(self.s), recs.ctypes.data_as(ctypes.c_void_p),
                                                    ctypes.c_size_t(recs.size)), "mpres_array_host2device")

    def host2device_ptr(self, ptr, count):
        _check(self.ctx.lib.mpres_array_host2device(self.ctx.h, ctypes.byref(self.s), ctypes.c_void_p(ptr),
                                                    ctypes.c_size_t(count)), "mpres_array_host2device")

    def host2device_ptr_at(self, offset, ptr, count):
        """`count` host records into elements offset .. offset + count (a shifted view of the same arrays; `len` still gives the
        offset of the upper interval bounds)"""
        n = self.ctx.N
        if offset < 0 or count < 0 or offset + count > max(1, self.size):
            raise ValueError("records %d .. %d do not fit an array of %d" % (offset, offset + count, self.size))
        v = mp_array_t(self.digits.data_ptr() + 4 * n * offset, self.sign.data_ptr() + 4 * offset, self.exp.data_ptr() + 4 * offset,
                       self.eval.data_ptr() + 16 * offset, None, self.len.data_ptr())
        _check(self.ctx.lib.mpres_array_host2device(self.ctx.h, ctypes.byref(v), ctypes.c_void_p(ptr), ctypes.c_size_t(count)),
               "mpres_array_host2device")

    def slices(self, offset, count):
        """the five contiguous device ranges that hold elements offset .. offset + count (digits, sign, exp, lower and upper bounds)"""
        from .parallel import soa_ranges
        return soa_ranges(self.digits, self.sign, self.exp, self.eval, self.ctx.N, max(1, self.size), offset, count)

    def device2host_ptr(self, ptr, count):
        _check(self.ctx.lib.mpres_array_device2host(self.ctx.h, ctypes.c_void_p(ptr), ctypes.byref(self.s),
                                                    ctypes.c_size_t(count)), "mpres_array_device2host")

    def device2host(self, size=None):
        import numpy as np
        n = self.size if size is None else int(size)
        out = np.zeros(n, dtype=self.ctx.dtype)
        self.device2host_ptr(out.ctypes.data, n)
        return out


def random_fill(ctx, arr, bits, seed, offset=0, count=None, chunk=1 << 22):
    """Fill arr[offset:offset+count] with synthetic values following tests/tsthelper.cuh:37-70:
    uniform `bits`-bit integer x uniform double in (-1, 1) x 2^-bits, rounded to `bits` bits.  The top 53
    bits and the exponent come from the double product computed in fp64, the lower bits are uniform
    random (statistically what the reference's MPFR product leaves there).  Deterministic in `seed`."""
    count = arr.size - offset if count is None else count
    dev = arr.digits.device
    gen = torch.Generator(device=dev)
    nl = (bits + 31) // 32
    done = 0
    while done < count:
        c = min(chunk, count - done)
        gen.manual_seed(seed * 1000003 + done)
        z = torch.rand(c, dtype=torch.float64, device=dev, generator=gen)          # Z / 2^bits in [0, 1)
        u = torch.rand(c, dtype=torch.float64, device=dev, generator=gen) * 2 - 1  # uniform (-1, 1)
        v = (z * u.abs()).clamp_min(2.0 ** -200)
        mant, e = torch.frexp(v)                                                   # v = mant * 2^e, mant in [0.5, 1)
        top = (mant * (1 << 53)).to(torch.int64)                                   # 53 significant bits
        rnd = torch.randint(0, 1 << 32, (c, nl), dtype=torch.int64, device=dev, generator=gen)
        # S = (top 53 bits << sh) | (sh uniform low bits), a normalised `bits`-bit significand
        sh = bits - 53
        cols = []
        for w in range(nl):
            lo_bit = 32 * w
            if sh >= 0:
                s = sh - lo_bit
                if s >= 32:
                    hi = torch.zeros_like(top)
                elif s >= 0:
                    hi = (top << s) & 0xFFFFFFFF
                else:
                    hi = (top >> (-s)) & 0xFFFFFFFF if -s < 63 else torch.zeros_like(top)
                keep = 0xFFFFFFFF if lo_bit + 32 <= sh else ((1 << (sh - lo_bit)) - 1 if lo_bit < sh else 0)
                cols.append(hi | (rnd[:, w] & keep))
            else:
                t = top >> (-sh)
                cols.append((t >> lo_bit) & 0xFFFFFFFF)
        limbs = torch.stack(cols, dim=1)
        sign = (u < 0).to(torch.int32)
        exp = (e.to(torch.int32) - bits).contiguous()
        l32 = torch.where(limbs >= (1 << 31), limbs - (1 << 32), limbs).to(torch.int32).contiguous()
        _check(ctx.lib.mpres_array_set_binary(ctx.h, ctypes.byref(arr.s), ctypes.c_size_t(offset + done),
                                              ctypes.c_void_p(sign.data_ptr()), ctypes.c_void_p(exp.data_ptr()),
                                              ctypes.c_void_p(l32.data_ptr()), nl, ctypes.c_size_t(c), None),
               "mpres_array_set_binary")
        torch.cuda.synchronize(dev)
        done += c
