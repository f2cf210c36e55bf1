"""PopDel profile binary format: writer (for fixtures) and a small reader (for tests).

Layout (little-endian), restated from the reference writer/reader
(`/root/reference/insert_histogram_popdel.h:128-205` header, `:298-328` index,
`/root/reference/popdel_profile/window_podel.h:161-197` window records,
`/root/reference/workflow_popdel.h:201-241` one gzip member per 10 kbp index region):

  "POPDEL\\x01" | u32 indexRegionSize | u32 numRegions | numRegions x u64 offsets |
  u32 numReadGroups | per RG { u32 len(name)+1 | name | 0 | u32 median | f64 stddev |
  u32 readLength | u32 histStart | u32 histEnd | (histEnd-histStart) x f64 counts } |
  u32 numContigs | per contig { u32 len(name)+1 | name | 0 | i32 length } | body

  body: per (contig, index region) one gzip member (or raw bytes) holding 256-bp window records
  u32 chrom | u32 beginPos | per RG { u32 n | n x { u8 posOffset | i32 isize - median } }.

The product-side decoder is the C++ one in `popdel_b200/csrc/profile_reader.cpp`; this module
only produces test inputs for it and for the reference binary.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Sequence, Tuple

import numpy as np

MAGIC = b"POPDEL\x01"
INDEX_REGION_SIZE = 10000
PROFILE_WINDOW = 256


def _gzip_member(raw: bytes) -> bytes:
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    return c.compress(raw) + c.flush()


def write_profile(path: str,
                  rg_meta: Sequence[dict],
                  contigs: Sequence[Tuple[str, int]],
                  records: Dict[int, List[Tuple[np.ndarray, np.ndarray]]],
                  compressed: bool = True) -> None:
    """rg_meta[r] = dict(name, median, stddev, read_length, hist_start, hist_end, hist_counts).
    records[chrom][r] = (pos uint32 sorted, isize int32) for read group r on contig `chrom`."""
    n_rg = len(rg_meta)
    n_regions = sum(l // INDEX_REGION_SIZE + 1 for _, l in contigs)
    head = bytearray()
    head += MAGIC
    head += struct.pack("<II", INDEX_REGION_SIZE, n_regions)
    index_pos = len(head)
    head += b"\0" * (8 * n_regions)
    head += struct.pack("<I", n_rg)
    for m in rg_meta:
        name = m["name"].encode()
        head += struct.pack("<I", len(name) + 1) + name + b"\0"
        head += struct.pack("<IdIII", int(m["median"]), float(m["stddev"]), int(m["read_length"]),
                            int(m["hist_start"]), int(m["hist_end"]))
        counts = np.asarray(m["hist_counts"], dtype="<f8")
        assert counts.size == m["hist_end"] - m["hist_start"]
        head += counts.tobytes()
    head += struct.pack("<I", len(contigs))
    for name, length in contigs:
        nb = name.encode()
        head += struct.pack("<I", len(nb) + 1) + nb + b"\0" + struct.pack("<i", int(length))

    body = bytearray()
    index: List[List[int]] = [[0] * (l // INDEX_REGION_SIZE + 1) for _, l in contigs]
    base = len(head)
    for chrom in sorted(records.keys()):
        per_rg = records[chrom]
        assert len(per_rg) == n_rg
        medians = [int(m["median"]) for m in rg_meta]
        # group records by 256-bp window
        wins = [np.asarray(p, dtype=np.int64) // PROFILE_WINDOW for p, _ in per_rg]
        all_w = np.unique(np.concatenate(wins)) if n_rg else np.zeros(0, dtype=np.int64)
        # per RG: boundaries of each window in the sorted pos array
        lo = [np.searchsorted(w, all_w, side="left") for w in wins]
        hi = [np.searchsorted(w, all_w, side="right") for w in wins]
        cur_region = -1
        block = bytearray()

        def flush_block():
            nonlocal block
            if block:
                body.extend(_gzip_member(bytes(block)) if compressed else bytes(block))
                block = bytearray()

        for k, w in enumerate(all_w):
            begin = int(w) * PROFILE_WINDOW
            region = begin // INDEX_REGION_SIZE
            if region != cur_region:
                flush_block()
                index[chrom][region] = base + len(body)
                cur_region = region
            block += struct.pack("<II", chrom, begin)
            for r in range(n_rg):
                a, b = int(lo[r][k]), int(hi[r][k])
                block += struct.pack("<I", b - a)
                if b > a:
                    rec = np.empty(b - a, dtype=np.dtype([("o", "u1"), ("d", "<i4")]))
                    rec["o"] = (per_rg[r][0][a:b].astype(np.int64) - begin).astype(np.uint8)
                    rec["d"] = (per_rg[r][1][a:b].astype(np.int64) - medians[r]).astype(np.int32)
                    block += rec.tobytes()
        flush_block()
    eof = base + len(body)
    # back-fill empty index entries with the next non-empty offset, trailing ones with EOF
    prev = eof
    for i in range(len(index) - 1, -1, -1):
        for j in range(len(index[i]) - 1, -1, -1):
            if index[i][j] == 0:
                index[i][j] = prev
            else:
                prev = index[i][j]
    flat = [o for c in index for o in c]
    head[index_pos:index_pos + 8 * n_regions] = struct.pack("<%dQ" % n_regions, *flat)
    with open(path, "wb") as fh:
        fh.write(bytes(head))
        fh.write(bytes(body))


def rg_meta_from(rg) -> dict:
    """Header fields of a `simulate.ReadGroupData`."""
    return dict(name=rg.spec.name, median=rg.median, stddev=rg.stddev, read_length=rg.spec.read_length,
                hist_start=rg.hist_start, hist_end=rg.hist_end, hist_counts=rg.hist_counts)


def write_cohort(dirpath: str, samples, contig: Tuple[str, int], compressed: bool = True,
                 extra_contigs: Sequence[Tuple[str, int]] = ()) -> List[str]:
    """One profile per sample of a single-contig cohort; returns the file paths."""
    import os
    os.makedirs(dirpath, exist_ok=True)
    paths = []
    contigs = [contig] + list(extra_contigs)
    for s in samples:
        meta = [rg_meta_from(rg) for rg in s.read_groups]
        recs = {0: [(rg.pos, rg.isize) for rg in s.read_groups]}
        p = os.path.join(dirpath, s.name + ".profile")
        write_profile(p, meta, contigs, recs, compressed)
        paths.append(p)
    return paths


def read_profile(path: str, compressed: bool = True):
    """Reads a whole profile; returns (rg_meta, contigs, records) in `write_profile`'s shapes
    (records hold isize deviations, not insert sizes: records[chrom][r] = (pos, dev))."""
    data = open(path, "rb").read()
    assert data[:7] == MAGIC, "bad magic"
    off = 7
    irs, n_regions = struct.unpack_from("<II", data, off)
    off += 8
    index = struct.unpack_from("<%dQ" % n_regions, data, off)
    off += 8 * n_regions
    (n_rg,) = struct.unpack_from("<I", data, off)
    off += 4
    meta = []
    for _ in range(n_rg):
        (nl,) = struct.unpack_from("<I", data, off)
        off += 4
        name = data[off:off + nl - 1].decode()
        off += nl
        median, stddev, rl, hs, he = struct.unpack_from("<IdIII", data, off)
        off += 24
        counts = np.frombuffer(data, dtype="<f8", count=he - hs, offset=off).copy()
        off += 8 * (he - hs)
        meta.append(dict(name=name, median=median, stddev=stddev, read_length=rl, hist_start=hs,
                         hist_end=he, hist_counts=counts))
    (n_contigs,) = struct.unpack_from("<I", data, off)
    off += 4
    contigs = []
    for _ in range(n_contigs):
        (nl,) = struct.unpack_from("<I", data, off)
        off += 4
        name = data[off:off + nl - 1].decode()
        off += nl
        (length,) = struct.unpack_from("<i", data, off)
        off += 4
        contigs.append((name, length))
    body = data[off:]
    if compressed:
        raw = bytearray()
        rest = body
        while rest:
            d = zlib.decompressobj(31)
            raw += d.decompress(rest)
            rest = d.unused_data
        body = bytes(raw)
    pos: Dict[int, List[list]] = {}
    dev: Dict[int, List[list]] = {}
    o = 0
    rec_dt = np.dtype([("o", "u1"), ("d", "<i4")])
    while o < len(body):
        chrom, begin = struct.unpack_from("<II", body, o)
        o += 8
        pl = pos.setdefault(chrom, [[] for _ in range(n_rg)])
        dl = dev.setdefault(chrom, [[] for _ in range(n_rg)])
        for r in range(n_rg):
            (n,) = struct.unpack_from("<I", body, o)
            o += 4
            rec = np.frombuffer(body, dtype=rec_dt, count=n, offset=o)
            o += 5 * n
            pl[r].append(rec["o"].astype(np.uint32) + np.uint32(begin))
            dl[r].append(rec["d"].astype(np.int32))
    records = {}
    for chrom in pos:
        records[chrom] = [(np.concatenate(pos[chrom][r]) if pos[chrom][r] else np.zeros(0, np.uint32),
                           np.concatenate(dev[chrom][r]) if dev[chrom][r] else np.zeros(0, np.int32))
                          for r in range(n_rg)]
    return meta, contigs, records, dict(index_region_size=irs, index=index)


def write_profile_single_rg(path: str, meta: dict, contigs: Sequence[Tuple[str, int]], chrom: int,
                            pos: np.ndarray, isize: np.ndarray, compressed: bool = True) -> None:
    """Vectorised writer for the common case of one read group on one contig (same bytes as write_profile)."""
    pos = np.asarray(pos, dtype=np.int64)
    n = pos.size
    n_regions = sum(l // INDEX_REGION_SIZE + 1 for _, l in contigs)
    head = bytearray()
    head += MAGIC
    head += struct.pack("<II", INDEX_REGION_SIZE, n_regions)
    index_pos = len(head)
    head += b"\0" * (8 * n_regions)
    head += struct.pack("<I", 1)
    name = meta["name"].encode()
    head += struct.pack("<I", len(name) + 1) + name + b"\0"
    head += struct.pack("<IdIII", int(meta["median"]), float(meta["stddev"]), int(meta["read_length"]),
                        int(meta["hist_start"]), int(meta["hist_end"]))
    head += np.asarray(meta["hist_counts"], dtype="<f8").tobytes()
    head += struct.pack("<I", len(contigs))
    for cname, length in contigs:
        nb = cname.encode()
        head += struct.pack("<I", len(nb) + 1) + nb + b"\0" + struct.pack("<i", int(length))
    base = len(head)
    win = pos // PROFILE_WINDOW
    first = np.flatnonzero(np.r_[True, win[1:] != win[:-1]]) if n else np.zeros(0, dtype=np.int64)
    counts = np.diff(np.r_[first, n])
    wbegin = win[first] * PROFILE_WINDOW
    wstart = np.r_[0, np.cumsum(12 + 5 * counts)]                 # byte offset of each window record
    body = np.zeros(int(wstart[-1]), dtype=np.uint8)
    hdr = np.zeros(first.size, dtype=np.dtype([("c", "<u4"), ("b", "<u4"), ("n", "<u4")]))
    hdr["c"], hdr["b"], hdr["n"] = chrom, wbegin, counts
    hb = hdr.view(np.uint8).reshape(-1, 12)
    for k in range(12):
        body[wstart[:-1] + k] = hb[:, k]
    rec = np.zeros(n, dtype=np.dtype([("o", "u1"), ("d", "<i4")]))
    wi = np.repeat(np.arange(first.size), counts)
    rec["o"] = (pos - wbegin[wi]).astype(np.uint8)
    rec["d"] = (np.asarray(isize, dtype=np.int64) - int(meta["median"])).astype(np.int32)
    rb = rec.view(np.uint8).reshape(-1, 5)
    roff = wstart[:-1][wi] + 12 + 5 * (np.arange(n) - first[wi])
    for k in range(5):
        body[roff + k] = rb[:, k]
    region = wbegin // INDEX_REGION_SIZE
    rfirst = np.flatnonzero(np.r_[True, region[1:] != region[:-1]]) if first.size else np.zeros(0, dtype=np.int64)
    bounds = np.r_[wstart[:-1][rfirst], wstart[-1]]
    index = [[0] * (l // INDEX_REGION_SIZE + 1) for _, l in contigs]
    out = bytearray()
    raw = body.tobytes()
    for k in range(rfirst.size):
        index[chrom][int(region[rfirst[k]])] = base + len(out)
        blk = raw[int(bounds[k]):int(bounds[k + 1])]
        out += _gzip_member(blk) if compressed else blk
    prev = base + len(out)
    for i in range(len(index) - 1, -1, -1):
        for j in range(len(index[i]) - 1, -1, -1):
            if index[i][j] == 0:
                index[i][j] = prev
            else:
                prev = index[i][j]
    head[index_pos:index_pos + 8 * n_regions] = struct.pack("<%dQ" % n_regions, *[o for c in index for o in c])
    with open(path, "wb") as fh:
        fh.write(bytes(head))
        fh.write(bytes(out))
