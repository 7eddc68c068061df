"""Runs tools/umma_probe.cu on cuda:0 and prints which shared-memory descriptor settings make
tcgen05.mma read an operand MN-major out of the K-major byte image of its transpose (DESIGN.md
par. 9.3: the MLP backward wants dh2^T, h1^T, W2^T, W1B^T without transposed copies).

    gpurun -- python tools/umma_probe.py        # writes gpurun_out/umma_probe.json

Operands are small integers (exact in bf16, exact fp32 sums), so a correct setting reproduces
A @ B^T bit for bit.  Not part of the product or of the test suite."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libumma_probe.so")


class ProbeParams(C.Structure):
    _fields_ = [("idesc", C.c_uint32), ("a_in_tmem", C.c_int), ("n_ksteps", C.c_int), ("N", C.c_int),
                ("a_bytes", C.c_uint32), ("b_bytes", C.c_uint32),
                ("a_lbo", C.c_uint32), ("a_sbo", C.c_uint32), ("a_kstep", C.c_uint32),
                ("b_lbo", C.c_uint32), ("b_sbo", C.c_uint32), ("b_kstep", C.c_uint32)]


def build():
    src = os.path.join(HERE, "umma_probe.cu")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                        "-Xcompiler", "-fPIC", "-shared", "-o", LIB, src], check=True, cwd=HERE, env=env)
    return C.CDLL(LIB)


def bf16_bits(x):
    b = np.ascontiguousarray(x, np.float32).view(np.uint32)
    assert not (b & 0xFFFF).any(), "values must be exact in bf16"
    return (b >> 16).astype(np.uint16)


def kmajor_image(X):
    """[R, K] -> bytes in the product's operand layout: 8x8 core matrices of 128 contiguous bytes,
    [row group][k chunk][8 rows][8 elements] (lrf_common.cuh::oper_offset)."""
    R, K = X.shape
    assert R % 8 == 0 and K % 8 == 0
    chunks = K // 8
    r = np.arange(R)[:, None]; k = np.arange(K)[None, :]
    off = (((r >> 3) * chunks + (k >> 3)) * 8 + (r & 7)) * 8 + (k & 7)
    img = np.zeros(R * K, np.uint16)
    img[off.reshape(-1)] = bf16_bits(X).reshape(-1)
    return img.view(np.uint8)


def tmem_rows(X):
    """[128, K] -> uint32 [128][K/2], low half = even k (the .ts A operand of the render kernel)."""
    b = bf16_bits(X).astype(np.uint32)
    return np.ascontiguousarray(b[:, 0::2] | (b[:, 1::2] << 16)).astype(np.uint32).view(np.uint8)


def idesc(N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((N >> 3) << 17) | ((128 >> 4) << 24)


def operand(X, mode):
    """-> (bytes, candidate (name, lbo, sbo, kstep) settings) for X [rows = M or N, K]."""
    R, K = X.shape
    if mode == "k":        # validated: K-adjacent core matrices 128 B apart, row groups chunks*128 B apart
        return kmajor_image(X), [("k-major", 128, (K // 8) * 128, 256)]
    if mode == "mn":       # image of the transpose [K][R]; read it back MN-major
        img = kmajor_image(np.ascontiguousarray(X.T))
        along_mn, along_k = 128, (R // 8) * 128
        return img, [("lbo=K-stride sbo=MN-stride", along_k, along_mn, 2 * along_k),
                     ("lbo=MN-stride sbo=K-stride", along_mn, along_k, 2 * along_k)]
    raise ValueError(mode)


def main():
    import torch
    if not torch.cuda.is_available():
        sys.exit("umma_probe needs a GPU (run it under gpurun)")
    lib = build()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    results = []
    out_dir = os.path.join(os.path.dirname(HERE), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)

    def dump():
        with open(os.path.join(out_dir, "umma_probe.json"), "w") as f:
            json.dump(results, f, indent=1)

    # most wanted first: a faulting setting poisons the CUDA context and ends the run, the JSON written
    # so far survives
    dead = False
    for K, N in ((128, 128), (64, 128), (128, 80), (128, 16), (128, 96)):
        A = rng.integers(-3, 4, (128, K)).astype(np.float32)
        B = rng.integers(-3, 4, (N, K)).astype(np.float32)
        want = A @ B.T
        n_read = (N + 31) // 32 * 32
        for a_mode, b_mode in (("k", "k"), ("tmem", "k"), ("k", "mn"), ("tmem", "mn"), ("mn", "k"), ("mn", "mn")):
            if dead:
                break
            if a_mode == "tmem":
                a_img, a_cands = tmem_rows(A), [("tmem", 0, 0, 0)]
            else:
                a_img, a_cands = operand(A, a_mode)
            b_img, b_cands = operand(B, b_mode)
            a_t = torch.from_numpy(a_img.copy()).to(dev); b_t = torch.from_numpy(b_img.copy()).to(dev)
            for an, albo, asbo, aks in a_cands:
                for bn, blbo, bsbo, bks in b_cands:
                    if dead:
                        break
                    p = ProbeParams(idesc(N, a_mode == "mn", b_mode == "mn"), int(a_mode == "tmem"), K // 16,
                                    n_read, a_img.size if a_mode != "tmem" else 0, b_img.size,
                                    albo, asbo, aks, blbo, bsbo, bks)
                    row = dict(K=K, N=N, A=a_mode, B=b_mode, a_desc=an, b_desc=bn)
                    try:
                        out = torch.full((128, n_read), float("nan"), device=dev)
                        rc = lib.umma_probe_run(C.byref(p), C.c_void_p(a_t.data_ptr()), C.c_void_p(b_t.data_ptr()),
                                                C.c_void_p(out.data_ptr()), None)
                        torch.cuda.synchronize()
                        got = out.cpu().numpy()[:, :N]
                        row.update(rc=rc, max_abs_err=float(np.abs(got - want).max()) if rc == 0 else None,
                                   n_wrong=int((got != want).sum()) if rc == 0 else None,
                                   timeout=bool(rc == 0 and got[0, 0] == -12345.0))
                    except RuntimeError as e:          # sticky CUDA error: nothing more can run in this process
                        row.update(rc=-99, max_abs_err=None, error=str(e).splitlines()[0][:200])
                        dead = True
                    results.append(row)
                    print(row, flush=True)
                    dump()
    ok = [r for r in results if r["max_abs_err"] == 0.0]
    print(f"{len(ok)} of {len(results)} settings exact")


if __name__ == "__main__":
    main()
