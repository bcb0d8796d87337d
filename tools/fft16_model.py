#!/usr/bin/env python
"""Index-level model of the 16-points-per-thread FFT proposed in DESIGN.md section 8 (next round): a 512-point transform
owned by ONE warp, one shared-memory exchange and one shuffle stage instead of two shared-memory exchanges and named
barriers.  numpy only; every array axis named `lane` is a lane of the warp, `reg` a register slot, so the three phases
below are exactly the data each thread holds.  Checked against numpy.fft.fft.

  phase 1   lane t holds x[t + 32 e], e < 16        radix-16 DFT over e, then the twiddle W512^(t k1)
  exchange  (t, k1) -> lane (k1, h), slot j with t = h + 2 j      (the one shared-memory round trip: 16 + 16 128-bit ops)
  phase 2   radix-16 DFT over j                      B[k1][h][q]
  shuffle   lanes (k1, 0) <-> (k1, 1) swap the halves q >= 8 / q < 8 they do not finish themselves (8 complex values)
  phase 3   X[k1 + 16 q + 256 r] = B0[q] + (-1)^r W32^q B1[q]      lane h finishes q in [8h, 8h + 8), r = 0, 1
"""
import numpy as np

N = 512
rng = np.random.default_rng(0)
x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
W = lambda n, p: np.exp(-2j * np.pi * p / n)  # noqa: E731

# phase 1: per lane t, registers e = 0..15
held = np.array([[x[t + 32 * e] for e in range(16)] for t in range(32)])          # [lane t][reg e]
A = np.array([[sum(held[t, e] * W(16, e * k1) for e in range(16)) for k1 in range(16)] for t in range(32)])
A *= np.array([[W(512, t * k1) for k1 in range(16)] for t in range(32)])           # [lane t][reg k1]

# exchange: lane (k1, h) = k1 + 16 h receives slot j <- A[t = h + 2 j][k1]
C = np.zeros((32, 16), complex)
for k1 in range(16):
    for h in range(2):
        for j in range(16):
            C[k1 + 16 * h, j] = A[h + 2 * j, k1]

# phase 2: radix-16 DFT over j inside each lane
B = np.array([[sum(C[lane, j] * W(16, j * q) for j in range(16)) for q in range(16)] for lane in range(32)])  # [lane][q]

# shuffle stage + phase 3: lane (k1, h) finishes q in [8h, 8h+8) for r = 0, 1; it needs the partner's B for those q
X = np.zeros(N, complex)
moved = 0
for k1 in range(16):
    for h in range(2):
        lane, partner = k1 + 16 * h, k1 + 16 * (1 - h)
        for q in range(8 * h, 8 * h + 8):
            b0 = B[lane, q] if h == 0 else B[partner, q]      # even-t half belongs to h = 0
            b1 = B[lane, q] if h == 1 else B[partner, q]
            moved += 1                                         # one complex value received per (lane, q)
            for r in range(2):
                X[k1 + 16 * q + 256 * r] = b0 + (-1) ** r * W(32, q) * b1

err = np.abs(X - np.fft.fft(x)).max() / np.abs(x).max()
print(f"max error vs numpy.fft: {err:.2e}; complex values moved by the shuffle stage per lane: {moved // 32}")
assert err < 1e-12
# store coalescing: for register slot (q - 8h, r) the lanes of one half-warp write 16 consecutive outputs
for r in range(2):
    for qi in range(8):
        ks = [k1 + 16 * (qi + 8 * h) + 256 * r for h in range(2) for k1 in range(16)]
        assert ks[:16] == list(range(ks[0], ks[0] + 16)) and ks[16:] == list(range(ks[16], ks[16] + 16))
print("each store instruction writes two runs of 16 consecutive outputs")

# DCT post-processing (K2) / pre-processing (K5) need Z[k] together with Z[N - k]: in this layout the partner of the
# value in lane (k1, h), slot (q, r) sits in lane ((16 - k1) % 16, 1 - h), slot (15 - q, 1 - r) -- except k1 = 0, where it
# is lane (0, h'), slot ((16 - q) % 16, ...) -- always inside the same warp, so the separation of the two real spectra
# is a shuffle as well and the transform needs ONE shared-memory exchange in total (128 wavefronts per 512 points
# against 320 today).
owner = {}
for k1 in range(16):
    for h in range(2):
        for q in range(8 * h, 8 * h + 8):
            for r in range(2):
                owner[k1 + 16 * q + 256 * r] = (k1, h, q, r)
for k in range(1, N // 2):
    k1, h, q, r = owner[k]
    p1, ph, pq, pr = owner[N - k]
    assert p1 == (16 - k1) % 16
    if k1 != 0:
        assert (ph, pq, pr) == (1 - h, 15 - q, 1 - r)
print("Z[k] and Z[N-k] always live in the same warp: lane (k1, h) <-> lane ((16 - k1) % 16, 1 - h)")

# The exchange the forward DCT (K2, the DCT producers of the fused kernel) needs after the FFT: lane L finishes the outputs
# k < 256, i.e. its registers o = 2 qi (r = 0), and needs Z[N - k] for each of them.  Source lane and source register:
out_index = lambda lane, o: (lane & 15) + 16 * ((o >> 1) + 8 * (lane >> 4)) + 256 * (o & 1)  # noqa: E731  (f16_out_index)
where = {out_index(lane, o): (lane, o) for lane in range(32) for o in range(16)}
rule_generic, rule_k1_zero = set(), []
for lane in range(32):
    k1, h = lane & 15, lane >> 4
    for qi in range(8):
        k = out_index(lane, 2 * qi)
        src_lane, src_o = where[(N - k) % N]
        if k1:
            assert src_lane == ((16 - k1) % 16) + 16 * (1 - h) and src_o == 2 * (7 - qi) + 1
            rule_generic.add("lane (k1, h), o = 2 qi  <-  lane (16 - k1, 1 - h), o = 2 (7 - qi) + 1")
        else:
            rule_k1_zero.append(((h, qi), (src_lane >> 4, src_o)))
print("Z[N-k] for the DCT separation, k1 != 0:", *rule_generic)
print("k1 = 0 lanes, (h, qi) <- (h', o'):", rule_k1_zero)
