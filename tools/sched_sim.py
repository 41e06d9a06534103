"""Tile-level model of the chained forward launch (gemm_chain_kernel_2cta with dep_signal / dep_wait).

Pure CPU, no GPU needed: it replays a per-pair unit schedule through the same pipeline the kernel has (in-order TMA
producer, one MMA issuer, two TMEM accumulator stages, in-order epilogue, row-block dependency counters) with per-tile
costs taken from the role timelines in profiles/r1_gemm_role_timelines.txt, and reports the makespan. It is the tool the
host-side schedule of `gemm_prepare_chain` (csrc/gemm_host.cu) was chosen with: compare `layered` (what four separate
launches do) with `pipelined` orders before spending GPU time.

    python tools/sched_sim.py [--pairs 64] [--frames 8192]
"""
from __future__ import annotations

import argparse
from dataclasses import dataclass
from typing import Dict, List, Tuple

# default.ini forward chain at B frames: (name, N, K, us per 64-deep k-block, epilogue us per unit)
LAYERS = [
    ("F1", 2048, 1024, 0.40, 3.2),
    ("F2", 512, 2048, 0.38, 3.2),
    ("F3", 2048, 256, 0.40, 3.5),
    ("F4", 1024, 2048, 0.48, 4.5),
]
FILL_US = 1.2      # dependency satisfied -> first operands landed
SIGNAL_US = 0.8    # epilogue done -> stores complete and the row-block counter visible to the consumer's producer
LAUNCH_US = 3.0    # per kernel launch: setup + tail that PDL does not hide
SCHED_MAX = 16


@dataclass(frozen=True)
class Unit:
    layer: int
    m: int
    n: int


def build_units(frames: int) -> Tuple[List[List[Unit]], int]:
    mb = frames // 256
    return [[Unit(i, m, n) for m in range(mb) for n in range(N // 256)] for i, (_, N, _, _, _) in enumerate(LAYERS)], mb


def order_layered(frames: int) -> List[Unit]:
    units, _ = build_units(frames)
    return [u for layer in units for u in layer]


def order_pipelined(frames: int, group: int, lag: int = 1) -> List[Unit]:
    """Row blocks in groups of `group`; layer i of group g is emitted at wavefront g + lag*i (deeper layers trail)."""
    units, mb = build_units(frames)
    ngroups = (mb + group - 1) // group
    out: List[Unit] = []
    for wave in range(ngroups + lag * (len(LAYERS) - 1)):
        for i in range(len(LAYERS) - 1, -1, -1):       # deeper layers first: they unblock nothing but free TMEM sooner
            g = wave - lag * i
            if 0 <= g < ngroups:
                out += [u for u in units[i] if g * group <= u.m < (g + 1) * group]
    return out


def cost(u: Unit) -> Tuple[float, float]:
    _, _, K, tkb, epi = LAYERS[u.layer]
    return (K // 64) * tkb, epi


def assign(order: List[Unit], pairs: int, how: str) -> List[List[Unit]]:
    """how = 'load': least accumulated cost (what gemm_prepare_chain did); 'sim': earliest simulated start."""
    lists: List[List[Unit]] = [[] for _ in range(pairs)]
    if how == "load":
        load = [0.0] * pairs
        for u in order:
            best = min((p for p in range(pairs) if len(lists[p]) < SCHED_MAX), key=lambda p: load[p])
            lists[best].append(u)
            load[best] += max(cost(u))
        return lists
    # list scheduling on the pipeline model itself
    st = SimState(pairs)
    for u in order:
        best, best_t = -1, 0.0
        for p in range(pairs):
            if len(lists[p]) >= SCHED_MAX:
                continue
            t = st.peek(p, u)
            if best < 0 or t < best_t - 1e-9:
                best, best_t = p, t
        st.commit(best, u)
        lists[best].append(u)
    return lists


class SimState:
    def __init__(self, pairs: int):
        self.mma_free = [0.0] * pairs
        self.epi_done: List[List[float]] = [[] for _ in range(pairs)]
        self.row_ready: Dict[Tuple[int, int], List[float]] = {}
        self.prod_free = [0.0] * pairs

    def dep_time(self, u: Unit) -> float:
        if u.layer == 0:
            return 0.0
        need = LAYERS[u.layer - 1][1] // 256
        done = self.row_ready.get((u.layer - 1, u.m), [])
        if len(done) < need:
            return float("inf")
        return max(done) + SIGNAL_US

    def peek(self, p: int, u: Unit) -> float:
        dep = max(self.dep_time(u), self.prod_free[p])
        acc = self.epi_done[p][-2] if len(self.epi_done[p]) >= 2 else 0.0
        start = max(self.mma_free[p], dep + FILL_US if u.layer else dep, acc)
        mma, epi = cost(u)
        last_epi = self.epi_done[p][-1] if self.epi_done[p] else 0.0
        return max(start + mma, last_epi) + epi

    def commit(self, p: int, u: Unit) -> float:
        dep = max(self.dep_time(u), self.prod_free[p])
        assert dep != float("inf"), "order is not topological"
        acc = self.epi_done[p][-2] if len(self.epi_done[p]) >= 2 else 0.0
        start = max(self.mma_free[p], dep + FILL_US if u.layer else dep, acc)
        mma, epi = cost(u)
        self.prod_free[p] = dep                      # in-order producer: later units cannot be requested before this
        self.mma_free[p] = start + mma
        last_epi = self.epi_done[p][-1] if self.epi_done[p] else 0.0
        done = max(start + mma, last_epi) + epi
        self.epi_done[p].append(done)
        self.row_ready.setdefault((u.layer, u.m), []).append(done)
        return done


def simulate(lists: List[List[Unit]]) -> float:
    """Event-driven replay: pairs advance in the order their next unit becomes runnable."""
    pairs = len(lists)
    st = SimState(pairs)
    pos = [0] * pairs
    remaining = sum(len(l) for l in lists)
    end = 0.0
    while remaining:
        best, best_t = -1, float("inf")
        for p in range(pairs):
            if pos[p] < len(lists[p]):
                t = st.peek(p, lists[p][pos[p]])
                if t < best_t:
                    best, best_t = p, t
        assert best >= 0, "deadlock: no runnable unit"
        end = max(end, st.commit(best, lists[best][pos[best]]))
        pos[best] += 1
        remaining -= 1
    return end + FILL_US


def separate_launches(frames: int, pairs: int) -> float:
    units, _ = build_units(frames)
    total = 0.0
    for i, layer in enumerate(units):
        saved = LAYERS[:]
        lists: List[List[Unit]] = [[] for _ in range(pairs)]
        for j, u in enumerate(layer):
            lists[j % pairs].append(Unit(0, u.m, u.n))
        LAYERS[0] = saved[i]
        total += simulate(lists) + LAUNCH_US
        LAYERS[:] = saved
    return total


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--frames", type=int, default=8192)
    a = ap.parse_args()
    print(f"four launches (model)            : {separate_launches(a.frames, a.pairs):7.1f} us")
    for pairs in (a.pairs, 74):
        lay = order_layered(a.frames)
        print(f"[{pairs} pairs] layered / load        : {simulate(assign(lay, pairs, 'load')) + LAUNCH_US:7.1f} us")
        for group in (2, 4, 8, 16):
            for lag in (1, 2):
                o = order_pipelined(a.frames, group, lag)
                for how in ("load", "sim"):
                    t = simulate(assign(o, pairs, how)) + LAUNCH_US
                    print(f"[{pairs} pairs] pipelined g={group:2d} lag={lag} / {how:4s}: {t:7.1f} us")


if __name__ == "__main__":
    main()
