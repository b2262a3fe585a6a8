"""Executable specification of the DEVICE algorithm (what yacrd_b200/csrc/pileup.cu computes per read),
in plain Python, so the crossing formulation can be fuzzed against the literal oracle on CPU.

Events: every interval (b, e) gives a begin event key 2*b+1 and an end event key 2*e, so that after an
ascending sort ends precede begins at equal positions (the heap pops `head <= begin` before the push,
stack.rs:72-81). depth = running (+1 begin / -1 end) sum. With threshold c:

  up-crossing    = begin event that lifts depth c -> c+1
  down-crossing  = end event that drops depth c+1 -> c

They alternate U0, D0, U1, D1, ... and the reference's cleaned gap list is exactly
  [(0, U0) if U0 != 0] ++ [(D_t, U_{t+1})] ++ [(D_last, len) if D_last != len]
or [(0, len) if len != 0] when depth never exceeds c (no crossing at all).
"""

U32 = 0xFFFFFFFF


def bad_part_by_crossings(ovls, length, coverage):
    ev = sorted([2 * b + 1 for b, _ in ovls] + [2 * e for _, e in ovls])
    gaps = []
    depth = 0
    last_down = None
    seen_up = False
    for key in ev:
        pos = key >> 1
        if key & 1:
            depth += 1
            if depth == coverage + 1:  # up-crossing
                if not seen_up:
                    if pos != 0:
                        gaps.append((0, pos))
                    seen_up = True
                else:
                    gaps.append((last_down, pos))
        else:
            depth -= 1
            if depth == coverage:  # down-crossing
                last_down = pos
    if not seen_up:
        if length != 0:
            gaps.append((0, length & U32))
    elif last_down != length:
        gaps.append((last_down, length & U32))
    return gaps


def classify(length, gaps, not_covered):
    """type_of_read on the device: same u32 sum and the same IEEE f64 divide + compare."""
    bad = 0
    for b, e in gaps:
        bad = (bad + e - b) & U32
    if length == 0:
        ratio = float("nan") if bad == 0 else float("inf")
    else:
        ratio = float(bad) / float(length)
    if ratio > not_covered:
        return 2
    for b, e in gaps:
        if b != 0 and e != (length & U32):
            return 1
    return 0
