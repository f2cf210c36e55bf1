"""Parser for the dump format shared by oracle/ref_harness.cpp and oracle/popdel_oracle.cpp."""
import gzip


def parse_dump(path):
    """Returns dict(windows=[(pos, ncalls, [(count, sum_dev, sum_pos)...])],
    segments=[dict(raw=[call...], merged=[call...])]); call = dict(ints=..., lr, freq, samples=[13 ints])."""
    opener = gzip.open if str(path).endswith(".gz") else open
    windows, segments = [], []
    cur_list = None
    with opener(path, "rt") as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            k = t[0]
            if k == "W":
                vals = list(map(int, t[3:]))
                windows.append((int(t[1]), int(t[2]), [tuple(vals[i:i + 3]) for i in range(0, len(vals), 3)]))
            elif k == "S":
                segments.append(dict(index=int(t[1]), raw=[], merged=[]))
                cur_list = segments[-1]["raw"]
            elif k == "M":
                cur_list = segments[-1]["merged"]
            elif k == "C":
                ints = [int(t[1]), int(t[2]), int(t[3])] + [int(x) for x in t[6:]]
                cur_list.append(dict(ints=ints, lr=float.fromhex(t[4]), freq=float.fromhex(t[5]), samples=[]))
            elif k == "G":
                cur_list[-1]["samples"].append([int(x) for x in t[1:]])
    return dict(windows=windows, segments=segments)
