"""Regenerates the golden fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  reference_data.npz  -- the reference's own data files (data/*.dat) as float64 arrays: input grids,
                         the four golden convolution outputs that dist/test.sh:39-61 diffs against, and the
                         four Y_l^m known-answer grids of dist/S2kitHowTo.pdf section 2.4.2.
  oracle_vectors.npz  -- outputs of the reference itself (oracle/_ref: its unmodified sources compiled here
                         against oracle/fftw_stub) on fixed inputs, for the cases no stored file pins.

/root/reference does not exist on the GPU box, so tests read only these .npz files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402

DATA = "/root/reference/data"


def main():
    oracle.build()
    ref = {}
    for name in sorted(os.listdir(DATA)):
        if name.endswith(".dat"):
            ref[name[:-4]] = np.loadtxt(os.path.join(DATA, name))
    np.savez_compressed(os.path.join(HERE, "reference_data.npz"), **ref)

    out = {}
    for bw in (16, 64):
        R = oracle.Oracle(bw, "ref")
        rc, ic = R.gen_coeffs(1000)
        out[f"coef_seed1000_bw{bw}_r"], out[f"coef_seed1000_bw{bw}_i"] = rc, ic
        for fmt, tag in ((0, "complex"), (1, "real")):
            rd, idt = R.inverse(rc, ic, fmt)
            out[f"inv_{tag}_bw{bw}_r"], out[f"inv_{tag}_bw{bw}_i"] = rd, idt
            fr, fi = R.forward(rd, idt, fmt)
            out[f"fwd_{tag}_bw{bw}_r"], out[f"fwd_{tag}_bw{bw}_i"] = fr, fi
        out[f"weights_bw{bw}"] = R.weights()
        for m in ((0, 1, 2, 7, 14, 15) if bw == 16 else (0, 1, 2, 31, 62, 63)):
            out[f"table_bw{bw}_m{m}"] = R.table(m)
        R.close()
    # config C1: forward transform of data/s64.dat (zero imaginary part), both formats
    bw = 64
    R = oracle.Oracle(bw, "ref")
    s = ref["s64"].reshape(2 * bw, 2 * bw)
    z = np.zeros_like(s)
    for fmt, tag in ((0, "complex"), (1, "real")):
        fr, fi = R.forward(s, z, fmt)
        out[f"s64_fwd_{tag}_r"], out[f"s64_fwd_{tag}_i"] = fr, fi
    zr, zi = R.zonal(ref["f64"].reshape(2 * bw, 2 * bw), z, 1)
    out["f64_zonal_r"] = zr
    # a fully complex coefficient set (independent negative orders), seed 12345
    rng = np.random.RandomState(12345)
    rc, ic = rng.uniform(-1, 1, bw * bw), rng.uniform(-1, 1, bw * bw)
    out["coef_full_bw64_r"], out["coef_full_bw64_i"] = rc, ic
    rd, idt = R.inverse(rc, ic, 0)
    out["inv_full_bw64_r"], out["inv_full_bw64_i"] = rd, idt
    fr, fi = R.forward(rd, idt, 0)
    out["fwd_full_bw64_r"], out["fwd_full_bw64_i"] = fr, fi
    R.close()
    # config C3 sample: function k=0 of the batch at bw 256 -- strided sample of the outputs
    bw = 256
    R = oracle.Oracle(bw, "ref")
    rc, ic = R.gen_coeffs(1000)
    rd, idt = R.inverse(rc, ic, 0)
    fr, fi = R.forward(rd, idt, 0)
    out["bw256_inv_sample_r"] = rd.ravel()[::257].copy()
    out["bw256_inv_sample_i"] = idt.ravel()[::257].copy()
    out["bw256_fwd_sample_r"] = fr[::61].copy()
    out["bw256_fwd_sample_i"] = fi[::61].copy()
    out["bw256_table_m1_head"] = R.table(1)[:4096]
    out["bw256_table_m200"] = R.table(200)
    R.close()
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    for f in ("reference_data.npz", "oracle_vectors.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
