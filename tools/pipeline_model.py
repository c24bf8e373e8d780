#!/usr/bin/env python
"""Discrete-event model of the forward kernel's steady state at head_dim 128 (clocks per 128x128 tile).

Calibrated on the -DFA_TRACE timeline of the shipped kernel (profiles/timeline_r01_full_4096_1thread_per_row.txt):
S seen -> S in registers 150, row max + stats 413, first 3/4 of the exponentials 940, last quarter 383, softmax ->
MMA-warp hop 150, tcgen05.commit -> softmax hop 250, one 128x128x128 GEMM 512 tensor clocks.

Two schedules are modelled:
  * "two_stage": the shipped kernel - two query tiles per CTA, each a serial chain softmax(j) -> P V(j) -> Q K^T(j+1)
    because P_s overwrites half of S_s; the MMA warp issues PV0 QK0 PV1 QK1 in order;
  * "pipelined": DESIGN section 8 item 1 - one query tile, one O accumulator, three S buffers, Q K^T issued three
    tiles ahead, two softmax groups taking alternate tiles and sharing the running row maximum (a group may start its
    exponentials once the other group has published the maximum of the previous tile).
The tensor pipe executes in issue order; an operation starts when the previous one has finished and its inputs are
ready. Output: clocks per KV tile and the fraction of that time the tensor pipe is busy.

This is a planning aid, not a measurement: `python tools/pipeline_model.py` prints both schedules.
"""
from dataclasses import dataclass


@dataclass
class Lat:
    ld: int = 150          # S seen -> S in registers (tcgen05.ld x128 + wait)
    rowmax: int = 413      # row max, lazy-rescale decision, stats published
    exp34: int = 940       # first 3/4 of the exponentials + P stores (early signal to the MMA warp)
    exp_last: int = 383    # last quarter
    hop_p: int = 150       # softmax arrive -> MMA warp sees it (mbarrier + tcgen05 fence)
    hop_s: int = 250       # tcgen05.commit -> softmax warps see S
    gemm: int = 512        # one 128x128x128 MMA block on the tensor pipe
    hop_max: int = 80      # shared-memory publish of the running maximum between the two softmax groups


def two_stage(n_tiles: int = 64, lat: Lat = Lat()):
    """Shipped schedule. Returns (clocks per iteration in the steady state, tensor-pipe busy fraction)."""
    pipe_free = 0.0
    s_seen = [[0.0] * (n_tiles + 1) for _ in range(2)]
    p34 = [[0.0] * n_tiles for _ in range(2)]
    pall = [[0.0] * n_tiles for _ in range(2)]
    sm_free = [0.0, 0.0]
    qk_end_of_iter = []

    def run(ready, dur):
        nonlocal pipe_free
        start = max(pipe_free, ready)
        pipe_free = start + dur
        return pipe_free

    def softmax(s, j):
        start = max(s_seen[s][j], sm_free[s])
        p34[s][j] = start + lat.ld + lat.rowmax + lat.exp34
        pall[s][j] = p34[s][j] + lat.exp_last
        sm_free[s] = pall[s][j]

    for it in range(n_tiles + 1):
        for s in range(2):
            if it > 0:
                j = it - 1
                softmax(s, j)                                      # its inputs (S_s(j)) were produced last iteration
                run(p34[s][j] + lat.hop_p, lat.gemm * 6 / 8)       # P V, k-steps 0..5 after the early signal
                run(pall[s][j] + lat.hop_p, lat.gemm * 2 / 8)      # k-steps 6..7 after the rest of P
            if it < n_tiles:
                end = run(0.0, lat.gemm)                           # S_s(it) = Q_s K^T (K tile assumed resident)
                s_seen[s][it] = end + lat.hop_s
        qk_end_of_iter.append(pipe_free)
    lo, hi = n_tiles // 4, 3 * n_tiles // 4
    period = (qk_end_of_iter[hi] - qk_end_of_iter[lo]) / (hi - lo)
    return period, 4 * lat.gemm / period


def pipelined(n_tiles: int = 64, lat: Lat = Lat(), s_buffers: int = 3):
    """One accumulator, `s_buffers` S buffers, two softmax groups on alternate tiles."""
    pipe_free = 0.0
    s_seen = [0.0] * n_tiles
    p34 = [0.0] * n_tiles
    pall = [0.0] * n_tiles
    max_pub = [0.0] * n_tiles
    pv_end = [0.0] * n_tiles
    grp_free = [0.0, 0.0]

    def run(ready, dur):
        nonlocal pipe_free
        start = max(pipe_free, ready)
        pipe_free = start + dur
        return pipe_free

    def qk(j):
        # buffer j % s_buffers was last read by P V(j - s_buffers), which precedes this op in issue order
        s_seen[j] = run(0.0, lat.gemm) + lat.hop_s

    def softmax(j):
        g = j & 1
        start = max(s_seen[j], grp_free[g])
        local_max = start + lat.ld + lat.rowmax                    # needs only S(j)
        prev = max_pub[j - 1] + lat.hop_max if j > 0 else 0.0      # the other group's maximum after tile j-1
        go = max(local_max, prev)
        max_pub[j] = go
        p34[j] = go + lat.exp34
        pall[j] = p34[j] + lat.exp_last
        grp_free[g] = pall[j]

    for j in range(min(s_buffers, n_tiles)):
        qk(j)
    for j in range(n_tiles):
        softmax(j)
        run(p34[j] + lat.hop_p, lat.gemm * 6 / 8)
        pv_end[j] = run(pall[j] + lat.hop_p, lat.gemm * 2 / 8)
        if j + s_buffers < n_tiles:
            qk(j + s_buffers)
    lo, hi = n_tiles // 4, 3 * n_tiles // 4
    period = (pv_end[hi] - pv_end[lo]) / (hi - lo)
    return period, 2 * lat.gemm / period


def main():
    lat = Lat()
    p, f = two_stage(lat=lat)
    print(f"two_stage (shipped): {p:7.0f} clocks per iteration (2 query tiles x 1 KV tile), tensor pipe busy {f:.3f}")
    for nb in (2, 3):
        p, f = pipelined(lat=lat, s_buffers=nb)
        print(f"pipelined, {nb} S buffers: {p:7.0f} clocks per KV tile (1 query tile), tensor pipe busy {f:.3f}")
    # sensitivity: a faster softmax (tuning), and a slower one - in the pipelined schedule both groups are busy ~90 % of
    # the time, so they share each sub-partition's issue slots and MUFU far more than the two stages do today
    for scale in (0.9, 0.8, 1.15, 1.3):
        l2 = Lat(rowmax=int(lat.rowmax * scale), exp34=int(lat.exp34 * scale), exp_last=int(lat.exp_last * scale))
        print(f"softmax x{scale}: two_stage {two_stage(lat=l2)[1]:.3f}, pipelined(3) {pipelined(lat=l2)[1]:.3f}")


if __name__ == "__main__":
    main()
