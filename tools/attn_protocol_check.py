"""Explicit-state model check of the attention kernel's mbarrier protocol (CPU only, pure Python).

The joint-attention kernel (dit4sr_b200/csrc/attn.cu, namespace v2) synchronises two MMA-issuing warps, the in-order
tensor pipe and two softmax warpgroups through mbarriers that are waited on by PARITY.  A parity wait can only tell
consecutive phases apart: it passes too early when the barrier is two phases behind the one meant, and blocks for ever
when the barrier is already two phases ahead.  Both happened on the GPU at the end of round 1 (a softmax variant whose
warps may take a 2x longer path corrupted rows and then deadlocked; guard waits on older phases deadlocked at once), and
neither is visible to a test that happens to run with well-behaved timing.  This tool explores EVERY interleaving of
the agents for small tile counts and reports

  * premature passes (a wait succeeds although the phase it stands for has not completed),
  * data hazards (scores read before they are written / after they were overwritten, P V issued on incomplete P,
    O rescaled while P V of the previous step is still outstanding),
  * deadlocks (some agent not finished, nobody can move),

for the barrier schemes of attn_fwd2_kernel:

  "v2"  the default instance: s_full / p_full indexed by score buffer (slot % 3) and therefore used alternately by the two
        query tiles (every waiter sees every other phase); o_done[t] waited only when a row's reference max grew;
  "v5"  the SAFE instance (DIT4SR_ATTN_XREG=2; first built in tools/experiments/attn_v5_*.cu.txt): s_full / p_full
        indexed by slot % 6 = (buffer, tile), o_done[t] waited in every step before P is published;
  "six_barriers_only", "o_done_always_only": one ingredient each.

Timing assumptions can be switched on to find the envelope a scheme relies on (`mma_keeps_up`, `slots_in_order`,
`max_lead`, see check()).  Result (tests/test_attn_protocol.py):

  * "v5" is safe for EVERY interleaving - no timing assumption at all;
  * "v2" is safe exactly when (A1) the MMA side keeps up - P V of step j-1 has completed when P of step j is published -
    and (A2) P completes in slot order (the two query tiles alternate).  Drop A1 and a softmax warp passes an o_done
    wait one phase early (reads / rescales O under a running P V); drop A2 and an issuing warp passes p_full one phase
    early (P V on a buffer that still holds the other tile's scores).  Six barriers remove the need for A2, the
    unconditional o_done wait the need for A1.  On the GPU both hold by a wide margin for the shipped kernel (every
    warp does the same work per step; ~300 cycles of MMA work against a ~2800-cycle step), which is why 120/120 stress
    runs and every parity test pass - and why the speculative-softmax variant, whose warps may take a 2x longer path,
    broke A2 and needs the "v5" scheme.

Model: K/V tiles are always resident (the TMA ring is a standard full/empty pipeline and is not part of the hazard);
tcgen05.commit arrives when the tensor pipe, which executes the two issuers' MMAs in issue order, reaches it.
"""
import sys
from collections import deque

NBUF = 3


class Violation(Exception):
    pass


def _programs(n_tiles, protocol):
    """Static step lists of the two issuers (what attn.cu's issuer loop does, K/V waits removed)."""
    slots = 2 * n_tiles
    progs = []
    for my in (0, 1):
        p = []
        for n in range(min(NBUF, slots)):
            if ((n + 1) & 1) == my:
                p.append(("qk", n))
        for n in range(my, slots, 2):
            p.append(("wait_p", n))
            p.append(("pv", n))
            if n + NBUF < slots:
                p.append(("qk", n + NBUF))
        progs.append(tuple(p))
    return progs


def check(n_tiles, warps_per_group=2, protocol="v2", max_lead=None, mma_keeps_up=False, slots_in_order=False,
          max_states=3_000_000):
    """Returns (ok, message, states_explored).

    max_lead:      timing assumption - no softmax warp starts a tile-step more than this many steps ahead of the slowest
                   softmax warp of the CTA (None: any skew).
    mma_keeps_up:  timing assumption - P V of step j-1 has always completed by the time a softmax warp publishes P of
                   step j (on the GPU: ~300 cycles of issue + MMA against a ~2800-cycle tile-step).  The v5 protocol
                   turns exactly this assumption into an explicit wait.
    slots_in_order: timing assumption - P of slot n is complete before any warp publishes P of slot n+1 (the two query
                   tiles alternate; on the GPU both warpgroups do equal work and their scores arrive in slot order)."""
    slots = 2 * n_tiles
    W = warps_per_group
    # protocol = (number of s_full / p_full barriers, when o_done is waited): "v2" = (3, only on growth),
    # "v5" = (6, every step); the two mixed forms show which ingredient buys what
    nbar, o_always = {"v2": (NBUF, False), "v5": (2 * NBUF, True), "six_barriers_only": (2 * NBUF, False),
                      "o_done_always_only": (NBUF, True)}[protocol]

    def bidx(n):
        return n % nbar

    def bphase(n):
        return n // nbar

    issuer_prog = _programs(n_tiles, protocol)
    # barrier layout inside the state tuple: s_full[nbar], p_full[nbar], o_done[2]; each (phase, pending)
    S0, P0, O0 = 0, nbar, 2 * nbar
    counts = [1] * nbar + [W] * nbar + [1, 1]

    # softmax warp micro-steps per tile-step: 0 = wait s_full + load scores, 1 = (optional) wait o_done, 2 = store P + arrive
    # warp state: (j, micro, grow) ; finished: j == n_tiles, micro 0 = final o_done wait pending, 1 = done
    init_bars = tuple((0, c) for c in counts)
    init = (
        (0, 0),                                  # issuer pcs
        tuple((0, 0, 0) for _ in range(2 * W)),  # softmax warps, index = t * W + w
        init_bars,
        (),                                      # tensor pipe FIFO of ("QK"|"PV", n)
        (None,) * NBUF,                          # score buffer content: slot whose S it holds
        (0,) * NBUF,                             # bitmask of warps that stored P for that slot
        0,                                       # bitmask of executed P V (by slot)
        0,                                       # bitmask of slots whose P is complete (all warps stored)
    )

    def arrive(bars, i):
        ph, pend = bars[i]
        pend -= 1
        if pend == 0:
            ph, pend = ph + 1, counts[i]
        return bars[:i] + ((ph, pend),) + bars[i + 1:]

    def can_pass(bars, i, k, what):
        """Parity wait for logical phase k of barrier i: (passes?, raises on a premature pass)."""
        ph = bars[i][0]
        if (ph & 1) == (k & 1):
            return False           # same parity as the pending phase: blocks (correctly if ph == k, falsely if ph > k)
        if ph <= k:
            raise Violation(f"premature pass: {what} waits for phase {k} of barrier {i} but only {ph} phases completed")
        return True

    seen = {init}
    todo = deque([init])
    explored = 0
    while todo:
        st = todo.pop()
        explored += 1
        if explored > max_states:
            return False, f"state limit {max_states} reached", explored
        ipc, warps, bars, pipe, content, pst, pvdone, pdone = st
        succ = []
        try:
            # ---- issuers
            for my in (0, 1):
                pc = ipc[my]
                if pc >= len(issuer_prog[my]):
                    continue
                op, n = issuer_prog[my][pc]
                nipc = ipc[:my] + (pc + 1,) + ipc[my + 1:]
                if op == "wait_p":
                    if can_pass(bars, P0 + bidx(n), bphase(n), f"issuer {my} (P of slot {n})"):
                        succ.append((nipc, warps, bars, pipe, content, pst, pvdone, pdone))
                else:
                    succ.append((nipc, warps, bars, pipe + ((op, n),), content, pst, pvdone, pdone))
            # ---- tensor pipe (in order)
            if pipe:
                op, n = pipe[0]
                b = n % NBUF
                if op == "qk":
                    prev = content[b]
                    if prev is not None and not (pvdone >> prev) & 1:
                        raise Violation(f"data hazard: Q K^T of slot {n} overwrites buffer {b} before P V of slot {prev}")
                    nc = content[:b] + (n,) + content[b + 1:]
                    np_ = pst[:b] + (0,) + pst[b + 1:]
                    succ.append((ipc, warps, arrive(bars, S0 + bidx(n)), pipe[1:], nc, np_, pvdone, pdone))
                else:
                    if content[b] != n or pst[b] != (1 << W) - 1:
                        raise Violation(f"data hazard: P V of slot {n} runs on buffer {b} holding slot {content[b]} with P "
                                        f"from warps {pst[b]:b}")
                    succ.append((ipc, warps, arrive(bars, O0 + (n & 1)), pipe[1:], content, pst, pvdone | (1 << n), pdone))
            # ---- softmax warps
            min_j = min(w[0] for w in warps)
            for wi, (j, micro, grow) in enumerate(warps):
                t, w = divmod(wi, W)
                if j == n_tiles:
                    if micro == 0:   # epilogue: wait for the last P V
                        if can_pass(bars, O0 + t, n_tiles - 1, f"warp {wi} epilogue (o_done)"):
                            if not (pvdone >> (2 * (n_tiles - 1) + t)) & 1:
                                raise Violation(f"data hazard: warp {wi} reads O before the last P V")
                            succ.append((ipc, warps[:wi] + ((j, 1, 0),) + warps[wi + 1:], bars, pipe, content, pst, pvdone, pdone))
                    continue
                n = 2 * j + t
                b = n % NBUF
                if micro == 0:
                    if max_lead is not None and j - min_j > max_lead:
                        continue
                    if can_pass(bars, S0 + bidx(n), bphase(n), f"warp {wi} (S of slot {n})"):
                        if content[b] != n:
                            raise Violation(f"data hazard: warp {wi} loads slot {n} from buffer {b} holding slot {content[b]}")
                        choices = (0,) if j == 0 else ((1,) if o_always else (0, 1))
                        for g in choices:
                            succ.append((ipc, warps[:wi] + ((j, 1 if g else 2, g),) + warps[wi + 1:], bars, pipe, content,
                                         pst, pvdone, pdone))
                elif micro == 1:     # wait for P V of the previous step of this tile (O rescale / pacing)
                    if can_pass(bars, O0 + t, j - 1, f"warp {wi} step {j} (o_done)"):
                        if not (pvdone >> (n - 2)) & 1:
                            raise Violation(f"data hazard: warp {wi} rescales O of tile {t} before P V of slot {n - 2}")
                        succ.append((ipc, warps[:wi] + ((j, 2, 0),) + warps[wi + 1:], bars, pipe, content, pst, pvdone, pdone))
                else:                # store P, arrive
                    if mma_keeps_up and j > 0 and not (pvdone >> (n - 2)) & 1:
                        continue     # (assumption) the MMA side has not fallen a whole step behind
                    if slots_in_order and n > 0 and not (pdone >> (n - 1)) & 1:
                        continue     # (assumption) the other tile's previous slot is complete first
                    if content[b] != n:
                        raise Violation(f"data hazard: warp {wi} stores P of slot {n} into buffer {b} holding slot {content[b]}")
                    np_ = pst[:b] + (pst[b] | (1 << w),) + pst[b + 1:]
                    npd = pdone | (1 << n) if np_[b] == (1 << W) - 1 else pdone
                    succ.append((ipc, warps[:wi] + ((j + 1, 0, 0),) + warps[wi + 1:], arrive(bars, P0 + bidx(n)), pipe,
                                 content, np_, pvdone, npd))
        except Violation as e:
            return False, str(e), explored
        if not succ:
            done = all(ipc[m] >= len(issuer_prog[m]) for m in (0, 1)) and not pipe and all(
                w == (n_tiles, 1, 0) for w in warps)
            if not done:
                return False, (f"deadlock: issuers at {ipc} of {[len(p) for p in issuer_prog]}, warps {warps}, "
                               f"pipe {pipe}, barrier phases {[b[0] for b in bars]}"), explored
        for s in succ:
            if s not in seen:
                seen.add(s)
                todo.append(s)
    return True, "safe", explored


def main():
    for protocol, keeps_up, in_order in (("v2", False, False), ("v2", True, False), ("v2", True, True),
                                         ("six_barriers_only", True, False), ("o_done_always_only", False, True),
                                         ("v5", False, False)):
        for lead in (1, None):
            for n_tiles in (1, 2, 3, 4, 5, 6):
                ok, msg, n = check(n_tiles, 2, protocol, lead, keeps_up, in_order)
                print(f"{protocol} mma_keeps_up={keeps_up} slots_in_order={in_order} max_lead={lead} n_tiles={n_tiles}: "
                      f"{'OK  ' if ok else 'FAIL'} {n:8d} states  {msg}")
                if not ok:
                    break


if __name__ == "__main__":
    sys.exit(main())
