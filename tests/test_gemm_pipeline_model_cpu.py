"""A discrete-event model of the stage protocol of dgemm_sub_tma2_kernel (csrc/dgemm_tma.cuh): 8 warps with random
speeds, 4 stages, k-tile kt + 2 produced by warp (kt mod 8) at the top of its iteration kt after the "empty" mbarrier
of that stage, a "full" mbarrier per stage (parity waits), and the two places a stage can be released:

  * 'next_top' (the kernel): the arrive for k-tile kt - 1 sits behind the wait loop of k-tile kt, i.e. after every
    DMMA of kt - 1 has been issued and therefore after every fragment load of kt - 1 has returned;
  * 'early' (the first version): the arrive is issued while the last fragment loads of the k-tile are still in
    flight (ptxas had moved it in front of the last 24 DMMAs).

The model asserts the invariant the GPU run violated: no stage is rewritten while a load of its previous k-tile is
outstanding, and parity waits never alias (a waiter is at most one phase behind).  With 'early' and slow loads the
violation shows up; with 'next_top' it cannot."""
import heapq
import random

import pytest

STAGES, WARPS, DIST = 4, 8, 2


def simulate(KT, release, seed, load_latency):
    rng = random.Random(seed)
    full_phase = [0] * STAGES       # completed productions per stage
    empty_arrivals = [0] * STAGES   # arrivals of the current consumption phase
    empty_phase = [0] * STAGES      # completed consumption phases per stage
    reads_outstanding = [0] * KT    # fragment loads of k-tile j still in flight (over all warps)
    violations = []
    # every warp is a little state machine driven from a time-ordered queue
    q = []
    for w in range(WARPS):
        heapq.heappush(q, (rng.random(), w, "top", 0))
    full_phase[0] += 1  # k-tiles 0 and 1 are produced in the prologue
    if KT > 1:
        full_phase[1] += 1

    def arrive(kt):
        s = kt % STAGES
        empty_arrivals[s] += 1
        if empty_arrivals[s] == WARPS:
            empty_arrivals[s] = 0
            empty_phase[s] += 1

    def produce(now, j):
        # the TMA lands some time later; what matters is whether loads of the previous tenant are outstanding THEN
        land = now + 5.0 + rng.random()
        heapq.heappush(q, (land, -1, "land", j))

    while q:
        now, w, what, kt = heapq.heappop(q)
        if what == "land":
            prev = kt - STAGES
            if prev >= 0 and reads_outstanding[prev] > 0:
                violations.append(("stage rewritten under outstanding loads", kt, prev, reads_outstanding[prev]))
            full_phase[kt % STAGES] += 1
            continue
        if what == "loads_done":
            reads_outstanding[kt] -= 1
            continue
        if what == "top":
            if kt >= KT:
                continue
            s = kt % STAGES
            # full wait with parity (kt // STAGES) & 1: legal only if the barrier is in phase kt//STAGES or one later
            need = kt // STAGES + 1
            if full_phase[s] < need:
                heapq.heappush(q, (now + 0.2, w, "top", kt))  # spin
                continue
            if full_phase[s] > need:
                violations.append(("full barrier ran two phases ahead of a waiter", kt, w))
            if release == "next_top" and kt > 0:
                arrive(kt - 1)
            if w == kt % WARPS and kt + DIST < KT:
                nx = kt + DIST
                sp = nx % STAGES
                if nx >= STAGES and empty_phase[sp] < nx // STAGES:
                    heapq.heappush(q, (now + 0.2, w, "top_producer_wait", kt))
                    continue
                produce(now, nx)
            heapq.heappush(q, (now, w, "compute", kt))
            continue
        if what == "top_producer_wait":
            nx = kt + DIST
            sp = nx % STAGES
            if empty_phase[sp] < nx // STAGES:
                heapq.heappush(q, (now + 0.2, w, "top_producer_wait", kt))
                continue
            if empty_phase[sp] > nx // STAGES:
                violations.append(("empty barrier ran two phases ahead of the producer", kt, w))
            produce(now, nx)
            heapq.heappush(q, (now, w, "compute", kt))
            continue
        if what == "compute":
            # issue the fragment loads, then the DMMAs; the loads return `load_latency` later
            reads_outstanding[kt] += 1
            t_issue_last_load = now + 1.0 + rng.random() * 3.0 * (1 + (w % 3))
            lat = load_latency(rng)
            heapq.heappush(q, (t_issue_last_load + lat, w, "loads_done", kt))
            if release == "early":
                # the arrive leaves right behind the ISSUE of the last load
                heapq.heappush(q, (t_issue_last_load + 0.01, w, "early_arrive", kt))
            # the DMMAs that consume the last loads cannot issue before those have returned
            heapq.heappush(q, (t_issue_last_load + lat + 1.0, w, "top", kt + 1))
            continue
        if what == "early_arrive":
            arrive(kt)
            continue
    return violations


@pytest.mark.parametrize("seed", range(20))
def test_release_behind_the_next_wait_loop_never_rewrites_a_stage_under_loads(seed):
    slow_loads = lambda rng: 0.5 + (40.0 if rng.random() < 0.2 else 0.0)  # a contended LSU now and then
    assert simulate(64, "next_top", seed, slow_loads) == []


def test_early_release_is_caught_by_the_model():
    slow_loads = lambda rng: 0.5 + (40.0 if rng.random() < 0.2 else 0.0)
    hits = sum(bool(simulate(64, "early", seed, slow_loads)) for seed in range(20))
    assert hits > 0  # the hazard the LU exposed on the GPU: a TMA box lands while fragment loads are still queued
