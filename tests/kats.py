"""Known-answer vectors transcribed from the reference's own unit tests (inputs and expected outputs
only). Each entry cites the reference test that pins it."""

# (name, intervals, length, coverage, expected bad parts)  — src/stack.rs:312-390
STACK_KATS = [
    ("A", [(10, 990)], 1000, 0, [(0, 10), (990, 1000)]),                      # stack.rs:315-316,354-357
    ("B", [(10, 90)], 1000, 0, [(0, 10), (90, 1000)]),                        # stack.rs:318-319,358-361
    ("C", [(10, 490), (510, 990)], 1000, 0, [(0, 10), (490, 510), (990, 1000)]),  # stack.rs:321-323,362-365
    ("D", [(0, 990)], 1000, 0, [(990, 1000)]),                                # stack.rs:325-326,366
    ("E", [(10, 1000)], 1000, 0, [(0, 10)]),                                  # stack.rs:328-329,367
    ("F", [(0, 490), (510, 1000)], 1000, 0, [(490, 510)]),                    # stack.rs:331-333,368
    ("cov2", [(0, 425), (0, 450), (0, 475), (525, 1000), (550, 1000), (575, 1000)], 1000, 2,
     [(425, 575)]),                                                           # stack.rs:372-390
    # implied by the editor tests (bad parts that make the expected scrubb/split output):
    ("scrubb_keep_begin_end", [(0, 4), (9, 13), (18, 22)], 22, 0, [(4, 9), (13, 18)]),  # scrubbing.rs:271-283
    ("scrubb_keep_middle", [(4, 18)], 22, 0, [(0, 4), (18, 22)]),             # scrubbing.rs:298-308
    ("split", [(9, 13), (18, 22)], 22, 0, [(0, 9), (13, 18)]),                # split.rs:259-270
]

# (bad parts, length, not_covered, expected type) — src/editor/mod.rs:114-128
NOT_BAD, CHIMERIC, NOT_COVERED = 0, 1, 2
TYPE_KATS = [
    ([(0, 10), (990, 1000)], 1000, 0.8, NOT_BAD),
    ([(0, 10), (90, 1000)], 1000, 0.8, NOT_COVERED),
    ([(0, 10), (490, 510), (990, 1000)], 1000, 0.8, CHIMERIC),
    ([(990, 1000)], 1000, 0.8, NOT_BAD),
    ([(0, 10)], 1000, 0.8, NOT_BAD),
    ([(490, 510)], 1000, 0.8, CHIMERIC),
]

# report text forms — src/stack.rs:276-278,415
REPORT_KATS = [
    ("SRR8494940.65223", 2706, [(0, 1131), (2690, 2706)], 0.8,
     "NotBad\tSRR8494940.65223\t2706\t1131,0,1131;16,2690,2706"),
    ("SRR8494940.141626", 30116, [(0, 326), (2957, 30116)], 0.8,
     "NotCovered\tSRR8494940.141626\t30116\t326,0,326;27159,2957,30116"),
    ("SRR8494940.91655", 15691, [(0, 151), (7213, 11269), (15633, 15691)], 0.8,
     "Chimeric\tSRR8494940.91655\t15691\t151,0,151;4056,7213,11269;58,15633,15691"),
    ("perfect", 2706, [], 0.8, "NotBad\tperfect\t2706\t"),
]

# contract quirks — oracle-derived, each explained by the cited reference lines
QUIRK_KATS = [
    # abutting coverage yields a zero-length gap and makes the read Chimeric (stack.rs:73-85)
    ("abut", [(0, 10), (10, 20)], 20, 0, [(10, 10)], CHIMERIC),
    # depth never exceeds c => single (0,len) gap => NotCovered (stack.rs:86-88,107-113,127-129)
    ("never_covered", [(5, 50), (60, 90)], 100, 2, [(0, 100)], NOT_COVERED),
    # begin==0 is the "unset" sentinel: no head gap (stack.rs:84-88,107)
    ("zero_begin", [(0, 50)], 100, 0, [(50, 100)], NOT_BAD),
    # duplicates are kept and both count toward depth (fullmemory.rs:84)
    ("dups", [(10, 90), (10, 90)], 100, 1, [(0, 10), (90, 100)], NOT_BAD),
    # a read with a length but no interval (add_length only, fullmemory.rs:78-80)
    ("empty", [], 100, 0, [(0, 100)], NOT_COVERED),
]
