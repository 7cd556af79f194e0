"""GPU bring-up probe: runs the tcgen05 self-test kernel over a table of operand-layout /
descriptor hypotheses and prints which ones reproduce A @ B.T exactly.  Each case runs in its own
subprocess so a trapped kernel cannot poison the next one.

usage: python tools/umma_probe.py            (on a B200)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (mode, n, k, variant)
    "kmajor_n256_k128": ("k", 256, 128, "std"),
    "kmajor_n16_k64": ("k", 16, 64, "std"),
    "kmajor_n64_k64_lbo0": ("k", 64, 64, "lbo0"),
    "mnmajor_n256_k128": ("mn", 256, 128, "std"),
    "mnmajor_n64_k64": ("mn", 64, 64, "std"),
    "mnmajor_n256_k128_swapped": ("mn", 256, 128, "swap"),
    "amn_bk_n128_k64": ("amn_bk", 128, 64, "std"),
    # no-swizzle K-major core-matrix layout (8 rows x 16 B contiguous), used for the 16-column aux operand
    "kns_n256_k16": ("kns", 256, 16, "std"),
    "kns_n256_k64": ("kns", 256, 64, "std"),
    "kns_n16_k16": ("kns", 16, 16, "std"),
    "a_sw128_b_kns_n256_k64": ("a_sw_b_kns", 256, 64, "std"),
    # no-swizzle MN-major ("chunk-major" save layout): [mn/8][k][8 mn halves]; variants swap LBO/SBO
    "mnns_n256_k128_a": ("mnns", 256, 128, "a"),
    "mnns_n256_k128_b": ("mnns", 256, 128, "b"),
    "mnns_n16_k128_a": ("mnns", 16, 128, "a"),
    "mnns_n16_k128_b": ("mnns", 16, 128, "b"),
    # CTA pair (cta_group::2, M = 256): B rows split between the CTAs
    "pair_k_n256_k128": ("pair_k", 256, 128, "std"),
    "pair_k_n16_k64": ("pair_k", 16, 64, "std"),
    "pair_kns_n256_k16": ("pair_kns", 256, 16, "std"),
    "pair_kns_n16_k16": ("pair_kns", 16, 16, "std"),
}


def run_case(name):
    import ctypes
    import numpy as np
    import torch
    import spnerf_b200
    from spnerf_b200 import _cabi, slab

    mode, n, k, variant = CASES[name]
    rng = np.random.default_rng(7)
    pair = mode.startswith("pair")
    a = rng.integers(-4, 5, size=(256 if pair else 128, k)).astype(np.float16)
    b = rng.integers(-4, 5, size=(n, k)).astype(np.float16)
    want = a.astype(np.float32) @ b.astype(np.float32).T
    args = _cabi.UmmaSelftest()
    ksteps = k // 16

    def kmajor(x, rows):
        img = slab.pack_matrix(x)
        offs = [(s // 4) * rows * 128 + (s % 4) * 32 for s in range(ksteps)]
        lbo = 0 if variant == "lbo0" else 16
        return img, offs, slab.smem_desc_template(lbo, 1024)

    def mnmajor(x):
        # x: (rows=M or N, K) -> store transposed: slabs of (K rows, 64 cols of M/N)
        xt = np.ascontiguousarray(x.T)                    # (K, rows)
        img = slab.pack_matrix(xt)
        offs = [s * 2048 for s in range(ksteps)]
        slab_bytes = k * 128
        if variant == "swap":
            return img, offs, slab.smem_desc_template(1024, slab_bytes)
        return img, offs, slab.smem_desc_template(slab_bytes, 1024)

    def kns(x):
        # no swizzle: [row/8][k/8][8 rows][8 halves]; LBO = 128 (next core matrix along K), SBO = (K/8)*128
        rows, kk = x.shape
        img = np.ascontiguousarray(x.reshape(rows // 8, 8, kk // 8, 8).transpose(0, 2, 1, 3)).view(np.uint8).reshape(-1)
        offs = [s * 256 for s in range(ksteps)]
        return img, offs, slab.smem_desc_template(128, (kk // 8) * 128, swizzle=0)

    def mnns(x):
        # x: (rows = M or N, K).  element (mn, k) at (mn/8)*(K*16) + k*16 + (mn%8)*2
        rows, kk = x.shape
        img = np.ascontiguousarray(x.reshape(rows // 8, 8, kk).transpose(0, 2, 1)).view(np.uint8).reshape(-1)
        offs = [s * 256 for s in range(ksteps)]
        mn_stride, k_stride = kk * 16, 128
        if variant == "a":
            return img, offs, slab.smem_desc_template(mn_stride, k_stride, swizzle=0)
        return img, offs, slab.smem_desc_template(k_stride, mn_stride, swizzle=0)

    if mode == "mnns":
        a_img, a_off, a_t = mnns(a)
        b_img, b_off, b_t = mnns(b)
        idesc = slab.idesc_f16(128, n, 1, 1)
    elif pair:
        f = kmajor if mode == "pair_k" else (lambda x, rows=None: kns(x))
        if mode == "pair_k":
            a0_img, a_off, a_t = kmajor(a[:128], 128)
            a1_img, _, _ = kmajor(a[128:], 128)
            b0_img, b_off, b_t = kmajor(b[:n // 2], n // 2)
            b1_img, _, _ = kmajor(b[n // 2:], n // 2)
        else:
            a0_img, a_off, a_t = kns(a[:128])
            a1_img, _, _ = kns(a[128:])
            b0_img, b_off, b_t = kns(b[:n // 2])
            b1_img, _, _ = kns(b[n // 2:])
        a_img = np.concatenate([a0_img, a1_img])
        b_img = np.concatenate([b0_img, b1_img])
        idesc = slab.idesc_f16(256, n, 0, 0)
    elif mode == "kns":
        a_img, a_off, a_t = kns(a)
        b_img, b_off, b_t = kns(b)
        idesc = slab.idesc_f16(128, n, 0, 0)
    elif mode == "a_sw_b_kns":
        a_img, a_off, a_t = kmajor(a, 128)
        b_img, b_off, b_t = kns(b)
        idesc = slab.idesc_f16(128, n, 0, 0)
    elif mode == "k":
        a_img, a_off, a_t = kmajor(a, 128)
        b_img, b_off, b_t = kmajor(b, n)
        idesc = slab.idesc_f16(128, n, 0, 0)
    elif mode == "mn":
        a_img, a_off, a_t = mnmajor(a)
        b_img, b_off, b_t = mnmajor(b)
        idesc = slab.idesc_f16(128, n, 1, 1)
    else:
        a_img, a_off, a_t = mnmajor(a)
        b_img, b_off, b_t = kmajor(b, n)
        idesc = slab.idesc_f16(128, n, 1, 0)

    dev = torch.device("cuda:0")
    ta = torch.from_numpy(a_img.copy()).to(dev)
    tb = torch.from_numpy(b_img.copy()).to(dev)
    td = torch.full((256 if pair else 128, n), float("nan"), device=dev)
    args.a_img, args.b_img, args.d_out = ta.data_ptr(), tb.data_ptr(), td.data_ptr()
    args.a_bytes, args.b_bytes = (ta.numel() // 2, tb.numel() // 2) if pair else (ta.numel(), tb.numel())
    args.n, args.ksteps, args.idesc = n, ksteps, idesc
    args.a_desc_template, args.b_desc_template = a_t, b_t
    for i in range(ksteps):
        args.a_off[i] = a_off[i]
        args.b_off[i] = b_off[i]
    fn = _cabi.lib().spnerf_selftest_umma
    if pair:
        fn = _cabi.lib().spnerf_selftest_umma2
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.POINTER(_cabi.UmmaSelftest), ctypes.c_void_p]
    rc = fn(ctypes.byref(args), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = td.cpu().numpy()
    err = float(np.nanmax(np.abs(got - want))) if not np.isnan(got).all() else float("nan")
    ok = bool(rc == 0 and np.array_equal(got, want))
    print(json.dumps({"case": name, "rc": rc, "ok": ok, "max_abs_err": err,
                      "nan": int(np.isnan(got).sum()), "watchdog": int(_cabi.lib().spnerf_watchdog_code())}))
    return 0 if ok else 1


def main():
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
        sys.exit(run_case(sys.argv[1]))
    results = {}
    for name in CASES:
        if len(sys.argv) > 1 and not name.startswith(sys.argv[1]):
            continue
        try:
            p = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=120)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            results[name] = json.loads(line[-1]) if line else {"ok": False, "rc": p.returncode,
                                                               "stderr": p.stderr[-400:]}
        except subprocess.TimeoutExpired:
            results[name] = {"ok": False, "timeout": True}
        print(name, results[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "umma_probe.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
