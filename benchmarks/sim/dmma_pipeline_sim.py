#!/usr/bin/env python
"""Discrete-time model of the warp-specialised DMMA banded x banded kernel (csrc/gbmm_dmma.cu) on ONE SM.

Purpose: find out, without GPU time, which structural knob of the kernel bounds the DMMA pipe utilisation
(profiles/dmma_r01_notes.md: measured 59-68 %).  The model keeps exactly what the kernel does structurally:
64x64 C tiles in dense coordinates over the band parallelogram, KC-wide k-chunks through an NSTAGE ring guarded by
full/empty barriers, consumer warps owning WM x WN blocks of the tile, empty 8x8x4 blocks skipped, one DMMA = one
16-cycle slot of the warp's sub-core pipe (warp w -> sub-core w % 4), CTAS CTAs resident per SM, a CTA walking TPC row
tiles of one 64-column block.  Time unit = one DMMA slot (16 cycles).

    python benchmarks/sim/dmma_pipeline_sim.py [--l 128] [--kc 16] [--stages 3] [--ctas 4] [--wm 32 --wn 16] ...
"""
import argparse
import itertools


def build_cta_work(l, u, TM, TN, KC, WM, WN, tpc, by0, j0, n):
    """Per tile: list over chunks of per-warp DMMA counts, for the CTA (column block j0, row tiles by0..by0+tpc)."""
    kla = kua = l
    klb = kub = u
    klc, kuc = kla + klb, kua + kub
    I0 = max(0, (j0 - kuc) // TM)
    warps = [(wi, wj) for wj in range(0, TN, WN) for wi in range(0, TM, WM)]
    tiles = []
    for t in range(tpc):
        i0 = (I0 + by0 + t) * TM
        if i0 >= n or i0 > j0 + TN - 1 + klc:
            continue
        plo = max(i0 - kla, j0 - kub, 0)
        phi = min(i0 + TM - 1 + kua, j0 + TN - 1 + klb, n - 1)
        plo &= ~3
        if phi < plo:
            continue
        nch = (phi - plo + KC) // KC
        chunks = []
        for it in range(nch):
            row = []
            for (wi, wj) in warps:
                cnt = 0
                for ks in range(KC // 4):
                    k0 = plo + it * KC + 4 * ks
                    k1 = k0 + 3
                    na = 0
                    for x in range(WM // 8):
                        a0 = i0 + wi + 8 * x
                        a1 = a0 + 7
                        # A block (rows a0..a1, cols k0..k1) touches the band  -kua <= i-k <= kla
                        if a0 - k1 <= kla and k0 - a1 <= kua:
                            na += 1
                    nb = 0
                    for y in range(WN // 8):
                        b0 = j0 + wj + 8 * y
                        b1 = b0 + 7
                        if k0 - b1 <= klb and b0 - k1 <= kub:
                            nb += 1
                    cnt += na * nb
                row.append(cnt)
            chunks.append(row)
        tiles.append(chunks)
    return tiles, len(warps)


def simulate(args):
    l = u = args.l
    TM = TN = 64
    n = 1 << 16
    gy = (4 * l + TN - 1) // TM + 2
    groups = (gy + args.tpc - 1) // args.tpc
    # a stream of CTAs in launch order (column blocks far from the matrix corners): (j0, by0)
    cta_stream = [(64 * (200 + c), g * args.tpc) for c in range(args.cols) for g in range(groups)]
    work_cache = {}

    def cta_chunks(j0, by0):
        key = by0
        if key not in work_cache:
            tiles, nw = build_cta_work(l, u, TM, TN, args.kc, args.wm, args.wn, args.tpc, by0, 64 * 200, n)
            flat = [c for t in tiles for c in t]
            tile_end = set(itertools.accumulate(len(t) for t in tiles))
            work_cache[key] = (flat, nw, tile_end)
        return work_cache[key]

    class Cta:
        def __init__(self, j0, by0, now):
            self.chunks, self.nw, self.tile_end = cta_chunks(j0, by0)
            self.nch = len(self.chunks)
            self.arrive = [None] * self.nch            # time chunk data lands
            self.released = [0] * self.nch             # consumers done with chunk
            self.issued = 0
            # warp state: [chunk index, phase, remaining]; phase 0 = wait data, 1 = overhead, 2 = dmma, 3 = epilogue
            self.w = [[0, 0, 0] for _ in range(self.nw)]
            self.start = now + args.prologue
            self.done_warps = 0

        def producer(self, now):
            while self.issued < self.nch and now >= self.start:
                gi = self.issued
                if gi >= args.stages and self.released[gi - args.stages] < self.nw:
                    break
                self.arrive[gi] = now + args.latency
                self.issued += 1

    useful = 0
    busy = [0, 0, 0, 0]
    now = 0
    stream = iter(cta_stream)
    live = []
    for _ in range(args.ctas):
        nxt = next(stream, None)
        if nxt:
            live.append(Cta(*nxt, now))
    total_dmma = 0
    rr = [0, 0, 0, 0]
    while live:
        for c in live:
            c.producer(now)
        # each sub-core issues at most one DMMA this slot
        cand = [[] for _ in range(4)]
        for ci, c in enumerate(live):
            for wi, st in enumerate(c.w):
                gi, ph, rem = st
                if ph == 3:                        # epilogue of the tile just finished (no DMMA pipe use)
                    if rem > 0:
                        st[2] -= 1
                    else:
                        st[1] = 0
                    continue
                if gi >= c.nch:
                    continue
                if ph == 0:
                    if c.arrive[gi] is not None and now >= c.arrive[gi]:
                        st[1], st[2] = 1, args.overhead
                        ph, rem = 1, args.overhead
                if ph == 1:
                    if rem > 0:
                        st[2] -= 1
                        continue
                    cnt = c.chunks[gi][wi]
                    st[1], st[2] = 2, cnt
                    ph, rem = 2, cnt
                if ph == 2:
                    if rem > 0:
                        cand[wi % 4].append((ci, wi))
                    else:
                        c.released[gi] += 1
                        if (gi + 1) in c.tile_end:
                            st[0], st[1], st[2] = gi + 1, 3, args.epilogue
                        else:
                            st[0], st[1], st[2] = gi + 1, 0, 0
        for s in range(4):
            if cand[s]:
                pick = cand[s][rr[s] % len(cand[s])]
                rr[s] += 1
                live[pick[0]].w[pick[1]][2] -= 1
                busy[s] += 1
                total_dmma += 1
        # retire finished CTAs
        for c in list(live):
            if all(st[0] >= c.nch and st[1] != 3 for st in c.w):
                live.remove(c)
                nxt = next(stream, None)
                if nxt:
                    live.append(Cta(*nxt, now))
        now += 1
        if now > 5_000_000:
            raise RuntimeError('simulation did not terminate')
    util = sum(busy) / (4.0 * now)
    return util, now, total_dmma


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--l", type=int, default=128)
    ap.add_argument("--kc", type=int, default=16)
    ap.add_argument("--stages", type=int, default=3)
    ap.add_argument("--ctas", type=int, default=4)
    ap.add_argument("--wm", type=int, default=32)
    ap.add_argument("--wn", type=int, default=16)
    ap.add_argument("--tpc", type=int, default=5)
    ap.add_argument("--cols", type=int, default=12, help="column blocks simulated")
    ap.add_argument("--latency", type=int, default=60, help="global->smem latency in DMMA slots (16 clk)")
    ap.add_argument("--overhead", type=int, default=6, help="per warp-chunk non-DMMA time (slots)")
    ap.add_argument("--epilogue", type=int, default=30, help="per warp-tile epilogue time (slots)")
    ap.add_argument("--prologue", type=int, default=100, help="CTA start-up (slots)")
    a = ap.parse_args()
    util, t, nd = simulate(a)
    print(f"l={a.l} kc={a.kc} stages={a.stages} ctas={a.ctas} warp={a.wm}x{a.wn} tpc={a.tpc} lat={a.latency} ovh={a.overhead} "
          f"epi={a.epilogue}: DMMA pipe busy {100 * util:.1f} %  ({nd} DMMAs in {t} slots)")


if __name__ == "__main__":
    main()
