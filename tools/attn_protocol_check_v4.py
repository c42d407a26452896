"""Explicit-state model check of the mbarrier protocol of attn_fwd3_kernel / attn_fwd4_kernel (CPU only, pure Python).

One query tile (the two tiles of a v3 CTA are independent; a v4 CTA has one): an MMA-issuing warp, the in-order tensor
pipe and W softmax warps synchronise through s_full[2], p_full[2] (double-buffered score tiles), o_done and o_final.
mbarrier waits are PARITY waits: try_wait.parity(k & 1) on a barrier that has completed c phases passes iff
(c & 1) != (k & 1) - right when c == k + 1, PREMATURE when c == k - 1, k - 3, ... and blocked for ever when c == k + 2.

Issuer:   QK(0); QK(1); for j: wait p_full[j&1] (phase j>>1); PV(j) -> commits o_done (+ o_final behind the last); QK(j+2)
Softmax:  for j: wait s_full[j&1] (phase j>>1); [online form, when a row max grew: wait o_done (phase j-1), rescale O];
          store P(j) into the score buffer; arrive p_full[j&1].     Epilogue: wait, then read O.
Epilogue schemes:
  "o_final"        wait o_final (phase 0): what the kernels do.
  "o_done_parity"  wait o_done with the parity of phase n-1: what v3 first shipped.  A warp may reach the epilogue while
                   P V(n-2) is still outstanding (c == n - 2), the wait passes at once and O is read two P V early - found
                   on the GPU as 3 split-KV runs in 36 that were off by 5e-3..8e-3 (profiles/r02_attn_split_fix.log).
The tool explores EVERY interleaving for small tile counts and reports premature passes, data hazards (scores read
before written, P V on incomplete P, O touched under an outstanding P V, O read before the last P V) and deadlocks."""
import sys
from collections import deque


class Violation(Exception):
    pass


def check(n_tiles, warps=2, epilogue="o_final", online=True, max_states=2_000_000):
    """Returns (ok, message, states explored)."""
    W = warps
    prog = [("qk", 0)] + ([("qk", 1)] if n_tiles > 1 else [])
    for j in range(n_tiles):
        prog += [("wait_p", j), ("pv", j)]
        if j + 2 < n_tiles:
            prog.append(("qk", j + 2))
    prog = tuple(prog)

    def passes(c, k, what):
        ok = (c & 1) != (k & 1)
        if ok and c < k + 1:
            raise Violation(f"premature pass: {what} waits for phase {k}, barrier has completed {c}")
        return ok

    # state: pc, pipe (tuple of ops), s_done[2], p_state[2] = (completions, pending), o_done, o_final,
    #        buf[2] (slot whose S the buffer holds or None), pmask[2] (warps that stored P of the slot in the buffer),
    #        pv_done (bitmask), warps: tuple of (j, micro)   micro: 0 wait S, 1 maybe wait o_done, 2 store+arrive; j == n: 0 epi wait, 1 done
    init = (0, (), (0, 0), ((0, W), (0, W)), 0, 0, (None, None), (0, 0), 0, tuple((0, 0) for _ in range(W)))
    seen = {init}
    todo = deque([init])
    explored = 0
    try:
        while todo:
            st = todo.popleft()
            explored += 1
            if explored > max_states:
                return False, f"state budget exceeded ({max_states})", explored
            pc, pipe, s_done, p_st, o_done, o_final, buf, pmask, pv_done, ws = st
            succ = []
            # ---- tensor pipe executes its head
            if pipe:
                op, j = pipe[0]
                b = j & 1
                if op == "QK":
                    if buf[b] is not None and not (pv_done >> buf[b]) & 1:
                        raise Violation(f"QK({j}) overwrites buffer {b} before P V({buf[b]}) consumed it")
                    nb = list(buf); nb[b] = j
                    nm = list(pmask); nm[b] = 0
                    ns = list(s_done); ns[b] += 1
                    succ.append((pc, pipe[1:], tuple(ns), p_st, o_done, o_final, tuple(nb), tuple(nm), pv_done, ws))
                else:
                    if buf[b] != j or pmask[b] != (1 << W) - 1:
                        raise Violation(f"P V({j}) executes on incomplete P (buffer holds {buf[b]}, mask {pmask[b]:b})")
                    succ.append((pc, pipe[1:], s_done, p_st, o_done + 1, o_final + (1 if j == n_tiles - 1 else 0), buf, pmask,
                                 pv_done | (1 << j), ws))
            # ---- issuer
            if pc < len(prog):
                op, j = prog[pc]
                if op == "qk":
                    succ.append((pc + 1, pipe + (("QK", j),), s_done, p_st, o_done, o_final, buf, pmask, pv_done, ws))
                elif op == "pv":
                    succ.append((pc + 1, pipe + (("PV", j),), s_done, p_st, o_done, o_final, buf, pmask, pv_done, ws))
                else:
                    if passes(p_st[j & 1][0], j >> 1, f"issuer p_full[{j & 1}]"):
                        succ.append((pc + 1, pipe, s_done, p_st, o_done, o_final, buf, pmask, pv_done, ws))
            # ---- softmax warps
            for w, (j, micro) in enumerate(ws):
                def upd(nj, nmicro, **kw):
                    nws = ws[:w] + ((nj, nmicro),) + ws[w + 1:]
                    return (pc, pipe, s_done, kw.get("p_st", p_st), o_done, o_final, buf, kw.get("pmask", pmask), pv_done, nws)
                if j == n_tiles:
                    if micro == 0:
                        c, k = (o_final, 0) if epilogue == "o_final" else (o_done, n_tiles - 1)
                        if passes(c, k, f"warp {w} epilogue"):
                            if not (pv_done >> (n_tiles - 1)) & 1:
                                raise Violation(f"warp {w} reads O before P V({n_tiles - 1}) completed")
                            succ.append(upd(j, 1))
                    continue
                b = j & 1
                if micro == 0:
                    if passes(s_done[b], j >> 1, f"warp {w} s_full[{b}]"):
                        if buf[b] != j:
                            raise Violation(f"warp {w} reads scores of slot {buf[b]} for step {j}")
                        succ.append(upd(j, 2))                       # no row max grew
                        if online and j > 0:
                            succ.append(upd(j, 1))                   # some row max grew: O must be rescaled
                elif micro == 1:
                    if passes(o_done, j - 1, f"warp {w} o_done (rescale)"):
                        if not (pv_done >> (j - 1)) & 1:
                            raise Violation(f"warp {w} rescales O under an outstanding P V({j - 1})")
                        succ.append(upd(j, 2))
                else:
                    if buf[b] != j:
                        raise Violation(f"warp {w} stores P({j}) into a buffer holding slot {buf[b]}")
                    nm = list(pmask); nm[b] |= 1 << w
                    c, pend = p_st[b]
                    pend -= 1
                    if pend == 0:
                        c, pend = c + 1, W
                    np_ = list(p_st); np_[b] = (c, pend)
                    succ.append(upd(j + 1, 0, p_st=tuple(np_), pmask=tuple(nm)))
            done = pc == len(prog) and not pipe and all(x == (n_tiles, 1) for x in ws)
            if not succ and not done:
                return False, f"deadlock: issuer pc {pc}/{len(prog)}, pipe {pipe}, warps {ws}, o_done {o_done}", explored
            for s2 in succ:
                if s2 not in seen:
                    seen.add(s2)
                    todo.append(s2)
    except Violation as e:
        return False, str(e), explored
    return True, "safe for every interleaving", explored


if __name__ == "__main__":
    for scheme in ("o_final", "o_done_parity"):
        for n in range(1, 7):
            ok, msg, k = check(n, 2, scheme)
            print(f"{scheme:14s} n_tiles={n}: {'OK ' if ok else 'BAD'} {msg} ({k} states)")
    sys.exit(0)
