"""Host-side multi-GPU logic on CPU: byte-balanced line sharding and a world_size-2 gloo run in
which every rank decodes its shard with the CPU oracle and rank 0 checks that the gathered shards equal the
whole batch — the N>1 path of bench.py without a GPU."""
import os
import socket
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent


def test_shard_by_bytes_properties(native):
    data, offs = native.generate(native.FMT_RFC5424, 5, 50_000)
    for G in (1, 2, 3, 4, 8):
        b = native.shard_by_bytes(offs, G)
        assert b[0] == 0 and b[-1] == len(offs) - 1 and (np.diff(b) >= 0).all()
        sizes = np.array([int(offs[b[g + 1]]) - int(offs[b[g]]) for g in range(G)])
        assert sizes.sum() == int(offs[-1])
        assert sizes.max() - sizes.min() <= 2 * 1024 + 2  # within two max-length lines of each other
    # degenerate: fewer lines than shards, empty batch, one huge line
    tiny = np.array([0, 5, 9], dtype=np.int32)
    b = native.shard_by_bytes(tiny, 8)
    assert b[0] == 0 and b[-1] == 2 and (np.diff(b) >= 0).all()
    assert native.shard_by_bytes(np.array([0], dtype=np.int32), 4).tolist() == [0, 0, 0, 0, 0]
    skew = np.array([0, 1_000_000, 1_000_010, 1_000_020], dtype=np.int32)
    b = native.shard_by_bytes(skew, 2)
    assert b.tolist()[0] == 0 and b.tolist()[-1] == 3


def _worker(rank: int, world: int, port: int, out_path: str):
    import torch.distributed as dist
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / "oracle"))
    import flowgger_b200 as fb
    import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data, offs = fb.generate(fb.FMT_RFC5424, 42, 20_000)  # same batch on every rank (deterministic generator)
        bounds = fb.shard_by_bytes(offs, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        base = int(offs[lo])
        sub_offs = (offs[lo:hi + 1] - base).astype(np.int32)
        sub_data = np.ascontiguousarray(data[base:int(offs[hi])]) if hi > lo else np.zeros(1, np.uint8)
        buf, o = pyoracle.decode_dump(fb.FMT_RFC5424, sub_data, sub_offs, nthreads=2)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((lo, hi, buf), gathered, dst=0)
        if rank == 0:
            whole, _ = pyoracle.decode_dump(fb.FMT_RFC5424, data, offs, nthreads=2)
            parts = sorted(gathered)
            assert parts[0][0] == 0 and parts[-1][1] == len(offs) - 1
            assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
            ok = b"".join(p[2] for p in parts) == whole
            Path(out_path).write_text("ok" if ok else "mismatch")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gloo_world2_shards_reassemble(native, oracle, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_utf8_validator_matches_python(native):
    H = native.load_host()
    import ctypes
    cases = [b"", b"abc", "é€😀".encode(), b"\xc0\x80", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xe2\x82", b"\xff", b"a\x80",
             b"\xf0\x9f\x9a\x80", b"\xc2", b"\xe0\x80\x80", b"\xf0\x80\x80\x80", "﻿<13>".encode()]
    for c in cases:
        try:
            c.decode("utf-8")
            want = 1
        except UnicodeDecodeError:
            want = 0
        buf = ctypes.create_string_buffer(c, len(c) or 1)
        assert H.fgh_is_valid_utf8(buf, len(c)) == want, c
