"""Host-side clip sharding: partition properties and the N>1 gather over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distilcodec_nabeel_b200.sharding import gather_by_clip, run_sharded, shard_clips, shard_sizes


@pytest.mark.parametrize("n,ws", [(0, 1), (1, 1), (1, 8), (7, 2), (256, 8), (1000, 8), (5, 3)])
def test_partition_is_contiguous_balanced_and_complete(n, ws):
    shards = [shard_clips(n, ws, r) for r in range(ws)]
    flat = [i for s in shards for i in s]
    assert flat == list(range(n))
    sizes = [len(s) for s in shards]
    assert max(sizes) - min(sizes) <= 1 and sizes == shard_sizes(n, ws)


def test_bad_arguments():
    with pytest.raises(ValueError):
        shard_clips(4, 0, 0)
    with pytest.raises(ValueError):
        shard_clips(4, 2, 2)
    with pytest.raises(ValueError):
        shard_clips(-1, 2, 0)


def test_single_process_gather_is_identity():
    x = torch.arange(12).reshape(4, 3)
    assert gather_by_clip(x, 4) is x
    with pytest.raises(ValueError):
        gather_by_clip(x, 5)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        clips = [torch.full((3,), float(i)) for i in range(n_clips)]
        idx, res = run_sharded(lambda c: c * 2 + 1, clips, ws, rank)           # the per-clip "hot path"
        local = torch.stack(res) if res else torch.empty(0, 3)
        full = gather_by_clip(local, n_clips)
        only0 = gather_by_clip(local, n_clips, dst=0)
        codes = gather_by_clip(torch.tensor(idx, dtype=torch.int64).reshape(-1, 1), n_clips)
        q.put((rank, full.tolist(), None if only0 is None else only0.tolist(), codes.flatten().tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 8])
def test_world_size_2_gloo_gather_restores_input_order(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[2.0 * i + 1] * 3 for i in range(n_clips)]
    for rank, full, only0, codes in results:
        assert full == expect                      # unsharded result, bit for bit, on every rank
        assert codes == list(range(n_clips))
        assert (only0 == expect) if rank == 0 else (only0 is None)


def test_time_tiles_cover_the_clip_in_order():
    from distilcodec_nabeel_b200.sharding import time_tiles
    for T, tile, halo in ((10, 4, 3), (4100, 1000, 96), (7, 100, 5), (8192, 8192, 96), (1, 1, 0)):
        tiles = time_tiles(T, tile, halo)
        assert tiles[0][2] == 0 and tiles[-1][3] == T
        for (lo, hi, s, e), nxt in zip(tiles, tiles[1:] + [None]):
            assert 0 <= lo <= s < e <= hi <= T and s - lo <= halo and hi - e <= halo
            assert (lo == 0 or s - lo == halo) and (hi == T or hi - e == halo)
            if nxt:
                assert nxt[2] == e


def test_partition_and_tiles_properties_hypothesis():
    """Property tests (hypothesis): shards are contiguous, balanced and cover [0, n); time tiles cover [0, T) once,
    in order, with halos clipped at the clip ends only."""
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    from distilcodec_nabeel_b200.sharding import shard_clips, shard_sizes, time_tiles

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 5000), st.integers(1, 64))
    def shards(n, ws):
        parts = [shard_clips(n, ws, r) for r in range(ws)]
        assert [i for p in parts for i in p] == list(range(n))
        sizes = [len(p) for p in parts]
        assert sizes == shard_sizes(n, ws) and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 100000), st.integers(1, 20000), st.integers(0, 300))
    def tiles(T, tile, halo):
        ts = time_tiles(T, tile, halo)
        assert [x for (_, _, s, e) in ts for x in (s, e)][0] == 0 and ts[-1][3] == T
        assert all(a[3] == b[2] for a, b in zip(ts, ts[1:]))
        for lo, hi, s, e in ts:
            assert lo == max(0, s - halo) and hi == min(T, e + halo) and 0 < e - s <= tile

    shards()
    tiles()


def test_rows_pitch_collapses_strided_views_for_dma():
    """Host logic of dc_copy2d_async: a time tile / per-clip crop of a (B, C, T) or (B, T) tensor is `rows` runs of
    `width` bytes at a constant pitch; views that do not collapse are refused rather than silently staged."""
    import pytest
    from distilcodec_nabeel_b200.sharding import _rows_pitch
    m = torch.empty(4, 128, 1000)
    assert _rows_pitch(m) == (512, 4000, 4000)
    assert _rows_pitch(m[1:3, :, 10:200]) == (256, 760, 4000)          # mel[b0:b1, :, lo:hi]
    c = torch.empty(5, 300, dtype=torch.int64)
    assert _rows_pitch(c[2:4, 7:57]) == (2, 400, 2400)                 # codes[b0:b1, s:e]
    assert _rows_pitch(c[1:2, 7:57]) == (1, 400, 400)
    assert _rows_pitch(torch.empty(3, 50)[:, 5:9]) == (3, 16, 200)
    assert _rows_pitch(m[:, ::2, :]) == (256, 4000, 8000)              # every other row still is one constant pitch
    with pytest.raises(ValueError):
        _rows_pitch(m[:, :100, :])                                     # two different leading pitches
    with pytest.raises(ValueError):
        _rows_pitch(m.transpose(1, 2))                                 # last dimension not contiguous
