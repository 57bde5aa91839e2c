"""Host-side logic of the multi-GPU path (videoseal_b200/dist.py) on CPU with the gloo backend, world_size 2 and 3: frame
sharding aligned to step_size, ragged all-gather reassembly, sharded extract_message; plus the oracle's sharding invariance
(frames are independent units) which is what makes the no-collective data path legal."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videoseal_b200.dist import shard_bounds


def test_shard_bounds_properties():
    for F in [0, 1, 7, 10, 64, 512, 513]:
        for world in [1, 2, 3, 8]:
            for step in [1, 4, 8]:
                b = shard_bounds(F, world, step)
                assert len(b) == world and b[0][0] == 0 and b[-1][1] == F
                for (s0, e0), (s1, e1) in zip(b, b[1:]):
                    assert e0 == s1 and s0 <= e0
                for s, e in b:
                    assert s % step == 0 or s == F          # key-frame groups are never split
                sizes = [e - s for s, e in b]
                assert max(sizes) - min(sizes) < 2 * step  # balanced to within one group (the last group may be ragged)
    assert shard_bounds(512, 8, 4) == [(i * 64, (i + 1) * 64) for i in range(8)]   # BASELINE configs[2]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, F, step):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from videoseal_b200.dist import embed_detect_sharded, extract_message_sharded
        g = torch.Generator().manual_seed(0)
        frames = torch.rand(F, 3, 8, 8, generator=g)
        msgs = torch.randint(0, 2, (1, 16), generator=g)
        W = torch.randn(3 * 8 * 8, 17, generator=g)

        # stand-in for the GPU model: key frame of every step-group is added to all frames of the group (video `repeat` mode)
        def embed_fn(x, m, offset=[0]):
            key = x[::step].repeat_interleave(step, dim=0)[: x.shape[0]]
            return x * 0.5 + key * 0.25

        def detect_fn(x):
            return x.flatten(1) @ W

        ref_imgs = embed_fn(frames, msgs)
        ref_logits = detect_fn(ref_imgs)
        imgs, logits, (s, e) = embed_detect_sharded(embed_fn, detect_fn, frames, msgs, step, gather_frames=True)
        assert imgs.shape == frames.shape and torch.equal(imgs, ref_imgs)
        assert torch.allclose(logits, ref_logits, atol=1e-5)
        msg = extract_message_sharded(logits[s:e])
        assert torch.equal(msg, (ref_logits[:, 1:].mean(0) > 0).unsqueeze(0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,F,step", [(2, 10, 4), (2, 64, 4), (3, 5, 4), (2, 0 + 3, 1)])
def test_sharded_embed_detect_gloo(world, F, step):
    mp.spawn(_worker, args=(world, _free_port(), F, step), nprocs=world, join=True)


def test_oracle_frame_sharding_invariance():
    """the property the sharding relies on: embedding a clip in step-aligned shards == embedding it whole"""
    from oracle import restate
    from tests.util import load_card, apply_overrides
    card = apply_overrides(load_card("videoseal_1.0"), {"num_blocks": 1, "depths": [1, 1, 1, 1]})
    spec = restate.spec_from_card(card)
    orc = restate.OracleModel(spec, restate.synth_state_dict(spec, 3))
    orc.step_size, orc.chunk_size = 2, 4
    g = torch.Generator().manual_seed(1)
    vid = torch.rand(6, 3, 256, 256, generator=g)
    msgs = torch.randint(0, 2, (1, spec["nbits"]), generator=g)
    with torch.no_grad():
        whole = orc.embed(vid, msgs, is_video=True)["imgs_w"]
        (s0, e0), (s1, e1) = shard_bounds(6, 2, orc.step_size)
        parts = torch.cat([orc.embed(vid[s0:e0], msgs, is_video=True)["imgs_w"], orc.embed(vid[s1:e1], msgs, is_video=True)["imgs_w"]])
        assert (whole - parts).abs().max().item() == 0.0
        d_whole = orc.detect(whole, is_video=True)["preds"]
        d_parts = torch.cat([orc.detect(whole[s0:e0], is_video=True)["preds"], orc.detect(whole[s1:e1], is_video=True)["preds"]])
        assert (d_whole - d_parts).abs().max().item() < 1e-5


def test_subshard_bounds_cover_the_clip_in_frame_order():
    from videoseal_b200.dist import subshard_bounds
    F, world, step, S = 512, 8, 4, 2
    b = subshard_bounds(F, world, step, S)
    assert len(b) == world and all(len(x) == S for x in b)
    # the all-gather of segment s concatenates rank 0..world-1 ranges: that must be a contiguous run of frames in order
    pos = 0
    for s in range(S):
        for r in range(world):
            a, e = b[r][s]
            assert a == pos and (e - a) % step == 0 and e > a
            pos = e
    assert pos == F
    import pytest
    with pytest.raises(ValueError):
        subshard_bounds(500, 8, 4, 2)
