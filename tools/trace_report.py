"""prints the stage timeline recorded by DMX_DEBUG=8 / DMX_TRACE (clock64 stamps of block 0, second task)"""
import sys
for path in sys.argv[1:]:
    rows = [list(map(int, l.split())) for l in open(path)]
    main = [r for r in rows[:256] if r[0]]
    if not main:
        continue
    t0 = main[0][0]
    print(path, "stages", len(main))
    for it in range(62, 68):
        if it < len(main):
            r = main[it]
            print(it, "L wait %6d got %6d | T start %6d full %6d empty %6d end %6d | C start %6d got %6d" % tuple(
                x - t0 for x in (r[0], r[1], r[2], r[3], r[4], r[7], r[5], r[6])))
    n = min(len(main) - 1, 120)
    print("cycles/stage", (main[n][6] - main[20][6]) / (n - 20))
    for w, r in enumerate(rows[256:265]):
        print("consumer warp", w, "stage 64..67 (start wait, got):", [x - t0 for x in r])
