"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref, built
from /root/reference by oracle/Makefile).  Runs only in the build container; the outputs are committed.

    python tests/golden/make_golden.py

For every moduli set with a _ref binary: deterministic inputs (tsthelper-style, fixed seeds) converted
with the reference's own mp_set_mpfr, and the reference HOST results of mp_mul, mp_add, mp_round,
rns_eval_compute[_fast], a sequential dot and a small v1-semantics gemm; plus every constant of
rns_const_init / mp_const_init.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import constants, gen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def raw(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def main():
    import ctypes
    for N in sorted(oracle.moduli_sets()):
        if not oracle.have_ref(N):
            continue
        ref = oracle.RefLib(N)
        L = ref.lib
        p = L.ref_mp_precision()
        out = {}
        for tag, bits in (("q", p // 4), ("f", p)):
            s, m, e = gen.random_values(96, bits, 4242 + N + bits)
            recs = ref.set_ints(s, m, e, bits)
            x, y = recs[:48], recs[48:]
            out[tag + "_sign"] = np.array(s, dtype=np.int32)
            out[tag + "_exp"] = np.array(e, dtype=np.int64)
            out[tag + "_mant"] = np.array([int(v).to_bytes(256, "little") for v in m], dtype="S256")
            out[tag + "_recs"] = raw(recs)
            mul = ref.host_mul(x, y)
            out[tag + "_mul"] = raw(mul)
            out[tag + "_add"] = raw(ref.host_add(x, y))
            out[tag + "_add_mixed"] = raw(ref.host_add(mul, x))
            rb = (np.arange(48) * 7 % max(2, bits - 1) + 1).astype(np.int32)
            out[tag + "_round_bits"] = rb
            out[tag + "_round"] = raw(ref.host_round(x, rb))
            out[tag + "_eval"] = raw(ref.host_eval(mul))
            out[tag + "_eval_fast"] = raw(ref.host_eval(x, fast=True))
            out[tag + "_dot"] = raw(np.array([ref.host_dot(x, y)]))
            mm, nn, kk = 5, 4, 6
            A, B, C = recs[:mm * kk], recs[30:30 + kk * nn], recs[60:60 + mm * nn]
            al, be = recs[90:91], recs[91:92]
            Cg, _ = ref.host_gemm(mm, nn, kk, al, A, B, be, C)
            out[tag + "_gemm"] = raw(Cg)
        np.savez_compressed(os.path.join(HERE, "ref_host_N%d.npz" % N), **out)
        # constants straight from the reference
        Lg = L.ref_moduli_product_log2()

        def geti(fn, count):
            a = np.zeros(count, dtype=np.int32)
            getattr(L, fn)(a.ctypes.data_as(ctypes.c_void_p))
            return a.tolist()
        rd, ru = np.zeros(N), np.zeros(N)
        L.ref_get_recip(rd.ctypes.data_as(ctypes.c_void_p), ru.ctypes.data_as(ctypes.c_void_p))
        d, i = np.zeros(5), np.zeros(5, dtype=np.int64)
        L.ref_get_eval_consts(d.ctypes.data_as(ctypes.c_void_p), i.ctypes.data_as(ctypes.c_void_p))
        import hashlib
        pow2 = geti("ref_get_pow2", (Lg + 1) * N)
        cst = {"N": N, "log2M": Lg, "mp_precision": p, "mp_h": L.ref_mp_h(), "mp_j": L.ref_mp_j(),
               "moduli": geti("ref_get_moduli", N), "part_inverse": geti("ref_get_part_inverse", N),
               "pow2_sha256": hashlib.sha256(np.array(pow2, dtype=np.int32).tobytes()).hexdigest(),
               "m_pow2_residues": geti("ref_get_m_pow2_residues", 30),
               "mi_pow2_residues_sha256": hashlib.sha256(np.array(geti("ref_get_mi_pow2_residues", 30 * N), dtype=np.int32).tobytes()).hexdigest(),
               "pow2_inverse_sha256": hashlib.sha256(np.array(geti("ref_get_pow2_inverse", 30 * N), dtype=np.int32).tobytes()).hexdigest(),
               "mrc_mult_inv_sha256": hashlib.sha256(np.array(geti("ref_get_mrc_mult_inv", N * N), dtype=np.int32).tobytes()).hexdigest(),
               "recip_rd_bits": [constants.double_bits(v) for v in rd], "recip_ru_bits": [constants.double_bits(v) for v in ru],
               "eval_doubles_bits": [constants.double_bits(v) for v in d], "eval_ints": i.tolist()}
        json.dump(cst, open(os.path.join(HERE, "constants_N%d.json" % N), "w"))
        print("golden for N=%d written (p=%d)" % (N, p))


if __name__ == "__main__":
    main()
