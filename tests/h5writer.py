"""Test helper: a byte-level writer for the HDF5 subset the mini reader understands (superblock v0,
contiguous little-endian f8 / i4 datasets in the root group) -- no h5py / libhdf5 in this image.
Also builds a complete synthetic input.h5 (every dataset CLOUDSC2_ARRAY_STATE%LOAD reads that reaches
the kernels) from SourceColumns + Params."""
import struct
from pathlib import Path

import numpy as np


def write_h5(path, datasets, per_node=16):
    """Write a superblock-v0 HDF5 file with contiguous datasets in the root group, byte by byte
    (same on-disk structures as config-files/reference.h5: TREE/HEAP/SNOD, v1 object headers).
    `datasets`: name -> float64 or int32 ndarray.  Names are spread over symbol-table nodes of
    `per_node` entries under one B-tree node, like libhdf5 does for larger groups."""
    assert 1 <= per_node <= 32
    names = sorted(datasets)
    heap_data = b"\0" * 8
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        s = n.encode() + b"\0"
        heap_data += s + b"\0" * (-len(s) % 8)
    heap_data += b"\0" * 32
    UNDEF = 0xFFFFFFFFFFFFFFFF
    pos = 96                      # after superblock (56 bytes + 40-byte root entry)
    root_ohdr = pos; pos += 16 + 24
    btree = pos; pos += 24 + 8 * (2 * 16 + 1) + 8 * 2 * 16
    heap = pos; pos += 32
    heap_data_addr = pos; pos += len(heap_data)
    groups = [names[k:k + per_node] for k in range(0, len(names), per_node)]
    assert len(groups) <= 32
    snods = []
    for _ in groups:
        snods.append(pos); pos += 8 + 40 * 32
    ohdrs, raws = {}, {}
    for n in names:
        ohdrs[n] = pos; pos += 16 + 256
    for n in names:
        raws[n] = pos; pos += datasets[n].nbytes + (-datasets[n].nbytes % 8)
    eof = pos
    b = bytearray(eof)
    b[0:8] = b"\x89HDF\r\n\x1a\n"
    b[8:16] = bytes([0, 0, 0, 0, 0, 8, 8, 0])
    b[16:24] = struct.pack("<HHI", 16, 16, 0)
    b[24:56] = struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    b[56:96] = struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", btree, heap)
    # root object header: one symbol-table message
    b[root_ohdr:root_ohdr + 16] = struct.pack("<BBHII", 1, 0, 1, 1, 24) + b"\0" * 4
    b[root_ohdr + 16:root_ohdr + 40] = struct.pack("<HHBBBB", 0x11, 16, 0, 0, 0, 0) + struct.pack("<QQ", btree, heap)
    b[btree:btree + 24] = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(groups), UNDEF, UNDEF)
    q = btree + 24
    b[q:q + 8] = struct.pack("<Q", 0); q += 8                     # key 0
    for grp, snod in zip(groups, snods):
        b[q:q + 16] = struct.pack("<QQ", snod, name_off[grp[-1]]); q += 16   # child k, key k+1
    b[heap:heap + 32] = b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap_data), len(heap_data) - 32, heap_data_addr)
    b[heap_data_addr:heap_data_addr + len(heap_data)] = heap_data
    for grp, snod in zip(groups, snods):
        b[snod:snod + 8] = b"SNOD" + struct.pack("<BBH", 1, 0, len(grp))
        for k, n in enumerate(grp):
            e = snod + 8 + 40 * k
            b[e:e + 40] = struct.pack("<QQII", name_off[n], ohdrs[n], 0, 0) + b"\0" * 16
    for n in names:
        a = datasets[n]
        msgs = b""
        dims = b"".join(struct.pack("<Q", d) for d in a.shape)
        body = struct.pack("<BBBB", 1, a.ndim, 0, 0) + b"\0" * 4 + dims
        msgs += struct.pack("<HHBBBB", 1, len(body), 0, 0, 0, 0) + body
        if a.dtype == np.float64:
            body = struct.pack("<BBBBI", 0x11, 0x20, 0x3f, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        else:
            body = struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 4) + struct.pack("<HH", 0, 32)
        body += b"\0" * (-len(body) % 8)
        msgs += struct.pack("<HHBBBB", 3, len(body), 1, 0, 0, 0) + body
        body = struct.pack("<BB", 3, 1) + struct.pack("<QQ", raws[n], a.nbytes)
        body += b"\0" * (-len(body) % 8)
        msgs += struct.pack("<HHBBBB", 8, len(body), 0, 0, 0, 0) + body
        o = ohdrs[n]
        b[o:o + 16] = struct.pack("<BBHII", 1, 0, 3, 1, len(msgs)) + b"\0" * 4
        b[o + 16:o + 16 + len(msgs)] = msgs
        b[raws[n]:raws[n] + a.nbytes] = a.tobytes()
    Path(path).write_bytes(bytes(b))


PARAM_DATASETS = {  # input.h5 name -> member of cloudsc2_params (yomcst/yoethf/yoecldp/yoephli loaders)
    "RG": "rg", "RD": "rd", "RCPD": "rcpd", "RETV": "retv", "RLVTT": "rlvtt", "RLSTT": "rlstt",
    "RLMLT": "rlmlt", "RTT": "rtt", "R2ES": "r2es", "R3LES": "r3les", "R3IES": "r3ies",
    "R4LES": "r4les", "R4IES": "r4ies", "R5LES": "r5les", "R5IES": "r5ies", "R5ALVCP": "r5alvcp",
    "R5ALSCP": "r5alscp", "RALVDCP": "ralvdcp", "RALSDCP": "ralsdcp", "RTWAT": "rtwat",
    "RTICE": "rtice", "RTWAT_RTICE_R": "rtwat_rtice_r", "YRECLDP_RCLCRIT": "rclcrit",
    "YRECLDP_RKCONV": "rkconv", "YRECLDP_RLMIN": "rlmin", "YRECLDP_RPECONS": "rpecons",
    "YREPHLI_RLPTRC": "rlptrc"}


def input_h5_datasets(src, prm) -> dict:
    """src: SourceColumns (fields as (KLEV[,+1], KLON) / (NDIM, KLEV, KLON) arrays), prm: Params."""
    f = src.f
    d = {"KLON": np.array([src.klon], dtype=np.int32), "KLEV": np.array([src.klev], dtype=np.int32),
         "PTSPHY": np.array([src.ptsphy], dtype=np.float64)}
    for name in ("pt", "pq", "pap", "paph", "plu", "plude", "pmfu", "pmfd", "pa", "psupsat", "pclv"):
        d[name.upper()] = np.ascontiguousarray(f[name], dtype=np.float64)
    cml = np.ascontiguousarray(f["tend_cml"], dtype=np.float64)          # (8, KLEV, KLON)
    d["TENDENCY_CML_T"], d["TENDENCY_CML_A"], d["TENDENCY_CML_Q"] = cml[0].copy(), cml[1].copy(), cml[2].copy()
    d["TENDENCY_CML_CLD"] = cml[3:].copy()
    for h5name, member in PARAM_DATASETS.items():
        d[h5name] = np.array([getattr(prm, member)], dtype=np.float64)
    return d
