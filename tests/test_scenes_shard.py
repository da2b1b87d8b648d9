"""Host logic: workload generators and the multi-GPU decomposition, incl. a world_size-2 gloo run."""
import os

import numpy as np
import pytest

from wave_tool_b200 import api, scenes, shard


def test_seeded_waves_are_valid_and_reproducible():
    for W, cfg in [(8, 2), (64, 3), (256, 4), (32, 5)]:
        a, b = scenes.seeded_waves(W, cfg), scenes.seeded_waves(W, cfg)
        assert a.dtype == np.float32 and a.shape == (W, 6) and np.array_equal(a, b)
        assert abs(a[:, 0].sum() - 1.0) < 1e-5                       # sum(A) = 1
        assert np.all(np.abs(np.hypot(a[:, 4], a[:, 5]) - 1.0) < 1e-3)  # |D| = 1 +- 1e-3 (render-engine.h:152-154)
        assert np.all((a[:, 3] >= 0.2) & (a[:, 3] <= 1.0)) and np.all(a[:, :3] >= 0)
        lam = 2 * np.pi / a[:, 1]
        assert lam.min() >= 2 - 1e-3 and lam.max() <= 64 + 1e-2
    assert not np.array_equal(scenes.seeded_waves(8, 2), scenes.seeded_waves(8, 3))


def test_procedural_textures():
    hm = scenes.procedural_heightmap(64)
    assert hm.shape == (64, 64, 4) and hm.dtype == np.uint8 and np.all(hm[..., 3] == 255) and np.array_equal(hm[..., 0], hm[..., 1])
    assert np.array_equal(hm, scenes.procedural_heightmap(64))
    sky = scenes.procedural_skybox(8)
    assert sky.shape == (6, 8, 8, 4)
    # centre of the +x face looks along +x: rgb ~ (1, .5, .5); of the -y face along -y: (.5, 0, .5)
    assert sky[0, 3:5, 3:5, 0].min() > 250 and abs(int(sky[0, 4, 4, 1]) - 128) < 20
    assert sky[3, 3:5, 3:5, 1].max() < 5


def test_animation_params():
    p = scenes.animation_params(90)
    assert p.waveAnimationTimeInSeconds == np.float32(1.5) and abs(p.verticalBounceWavePhase - 0.5) < 1e-6
    frames = scenes.frames_for_animation(scenes.camera(), scenes.default_waves(), first=1, count=3, stride=2)
    assert [f.waveAnimationTimeInSeconds for f in frames] == [np.float32(k) / np.float32(60) for k in (1, 3, 5)]


def test_shards_partition_exactly():
    for n, world in [(1000, 1), (1000, 8), (1024, 4), (7, 8), (5, 2)]:
        got = sorted(f for r in range(world) for f in shard.frame_shard(n, r, world))
        assert got == list(range(n))
    for L, world in [(16384, 8), (513, 8), (4096, 3), (2, 2)]:
        spans = [shard.row_shard(L, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == L
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.row_shard(10, 2, 2)


def test_combine_checksums_wraps_mod_2_64():
    r = np.zeros(2, api.CHECKSUM_DTYPE)
    r["position_bits"] = [2 ** 64 - 5, 9]
    r["min_y"] = [-1.5, 0.25]; r["max_y"] = [0.5, 2.0]; r["vertices"] = [10, 20]
    c = shard.combine_checksums(r)
    assert c["position_bits"] == 4 and c["min_y"] == -1.5 and c["max_y"] == 2.0 and c["vertices"] == 30


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_frames = 10
        mine = shard.frame_shard(n_frames, rank, world)
        local = np.zeros(len(mine), api.CHECKSUM_DTYPE)
        local["position_bits"] = [1000 + f for f in mine]         # stand-in digests keyed by the global frame index
        local["min_y"] = [-float(f) for f in mine]; local["max_y"] = [float(f) for f in mine]
        local["vertices"] = 4
        allr = shard.all_gather_checksums(local)
        t = shard.max_over_ranks(1.0 + rank)
        # re-interleave: frame f lives at [f % world][f // world]
        ordered = np.array([allr[f % world][f // world] for f in range(n_frames)], api.CHECKSUM_DTYPE)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), ordered)
        assert t == float(world)
    finally:
        dist.destroy_process_group()


def test_frame_sharded_gather_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b)
    assert list(a["position_bits"]) == [1000 + f for f in range(10)]
    assert shard.combine_checksums(a)["vertices"] == 40


@pytest.mark.parametrize("L,R,mode", [(4096, 8, 1), (4096, 4, 0), (16384, 8, 1), (8192, 12, 1), (1024, 8, 1), (513, 8, 1), (513, 4, 0), (7, 8, 1), (32768, 8, 1)])
def test_canonical_layout_is_shard_invariant(L, R, mode):
    """The work items of a row shard are the chunks of the CANONICAL layout of the whole grid that meet the shard
    (wt_debug_layout, host only): chunks tile [0, L) exactly, every chunk keeps its full canonical extent in every
    shard — the property that makes recurrence-mode results bit-identical across shards, streams and devices
    (wt_ctx.h: canonical_layout) — and the 8 row shards of a box get equal work."""
    full = api.debug_layout(L, R, mode)
    order = np.argsort(full[:, 0])
    srt = full[order]
    assert srt[0, 0] == 0 and srt[-1, 1] == L and np.array_equal(srt[1:, 0], srt[:-1, 1])       # exact tiling, no overlap
    canon = {tuple(c) for c in full}
    rng = np.random.default_rng(L + R)
    cuts = sorted({0, L, *rng.integers(1, L, size=min(6, L - 1)).tolist()})
    for world in (2, 8):
        if L >= world:
            cuts_w = [shard.row_shard(L, r, world) for r in range(world)]
            sizes = []
            for b, e in cuts_w:
                part = api.debug_layout(L, R, mode, rows=(b, e))
                assert {tuple(c) for c in part} <= canon
                assert all(c[1] > b and c[0] < e for c in part)                                   # only chunks that meet the shard
                cover = sorted(part.tolist())
                assert cover[0][0] <= b and cover[-1][1] >= e
                # rows the kernel marches for this shard: per chunk, the steps of R rows from the one that holds row r0 - 1 to the
                # one that holds row r1 ([r0, r1) = the chunk clipped to the shard), on the chunk's own lattice (wt_kernels.cuh)
                marched = 0
                for c0, c1 in part.tolist():
                    r0, r1 = max(c0, b), min(c1, e)
                    first = c0 - 1 + ((r0 - c0) // R) * R
                    marched += ((r1 - first) // R + 1) * R
                sizes.append(marched)
            if L >= 4096:
                assert max(sizes) <= 1.05 * min(sizes), sizes                                     # balanced marching work per shard
    for b, e in zip(cuts[:-1], cuts[1:]):
        part = api.debug_layout(L, R, mode, rows=(b, e))
        assert {tuple(c) for c in part} <= canon
        got = sorted(part.tolist())
        assert got[0][0] <= b < got[0][1] and got[-1][0] < e <= got[-1][1]
        assert all(x[1] == y[0] for x, y in zip(got[:-1], got[1:]))
    # big items are handed out before the small ones (the tail of the launch is filled with small pieces)
    h = full[:, 1] - full[:, 0]
    if len(set(h.tolist())) > 1 and L >= 4096:
        first_small = int(np.argmax(h < h.max() - R))
        assert np.all(h[:first_small] >= h[first_small:].max())
