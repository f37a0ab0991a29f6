"""Worker of tests/test_multi_gpu.py, launched as one process per GPU by torch.distributed.run.  Every rank builds its
share of a tile DB, the ranks re-cut it into equal row ranges (dunk_shard_group_balance, NCCL inside the library) and
register their own frame batches against the sharded DB; rank 0 repeats everything unsharded on its GPU and compares
record for record (the merge is exact: (distance, index) lexicographic, so the results must be IDENTICAL)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import cubesat_apds_b200 as dunk
    import synthdata
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fd = dunk.feature_database
    ctx = dunk.Context(local, 4)
    uid = [fd.ShardGroup.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    group = fd.ShardGroup(ctx, rank, world, uid[0])
    S = 2048
    scene = synthdata.synth_scene(S, seed=11)
    tiles = np.stack([scene[r * 512:(r + 1) * 512, c * 512:(c + 1) * 512] for r in range(4) for c in range(4)])
    xo = np.array([c * 512 for r in range(4) for c in range(4)], np.float32)
    yo = np.array([r * 512 for r in range(4) for c in range(4)], np.float32)
    lo, hi = 16 * rank // world, 16 * (rank + 1) // world           # contiguous share: global row order = rank-major
    built = fd.DescriptorDatabase(ctx, capacity=200000)
    built.append_tiles(tiles[lo:hi], xo[lo:hi], yo[lo:hi], 1.0, np.arange(lo, hi, dtype=np.int32) + 1)
    shard = group.balance(built)
    sizes = [group.base(r + 1) - group.base(r) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1 and len(shard) == sizes[rank]
    Hs, Rs, ts, _ = synthdata.config5_views(3, S, 100 + rank)
    hg = dunk.homographier
    frames = np.stack([hg.warp_image_perspective(hg.Cmat(scene, np.uint8), H, (1024, 1024), ctx).mat for H in Hs])
    if rank == world - 1:
        frames[1] = 77                                             # a flat frame: unequal query counts across ranks
    gt_e, heights = synthdata.scene_dem(S)
    geo = fd.Geotransform(synthdata.scene_geotransform(), gt_e, heights, ctx)
    pose = fd.PoseStage(geo, synthdata.CAMERA_K, synthdata.scene_origin(S), 500, 3.0, 0.99, 1)
    res, poses = group.register_frames(shard, frames, pose=pose)
    # config 3 form: the same queries on every rank against the sharded DB
    q = np.random.default_rng(5).integers(0, 256, (700, 61), dtype=np.uint8)
    mine = shard.read_descriptors(0, min(50, len(shard)))
    planted = [None] * world
    dist.all_gather_object(planted, mine)
    q = np.concatenate([q] + planted)
    m = group.match(shard, q, 0.85, group.base(rank))
    gathered = [None] * world
    dist.gather_object((frames, res, poses, m), gathered if rank == 0 else None, 0)
    ok = True
    if rank == 0:
        whole = fd.DescriptorDatabase(ctx, capacity=400000)
        whole.append_tiles(tiles, xo, yo, 1.0, np.arange(16, dtype=np.int32) + 1)
        assert len(whole) == group.total_rows
        d, _, _ = whole.read_rows(group.base(0), len(shard))
        assert np.array_equal(d, shard.read_descriptors(0, len(shard)))
        m_ref = whole.match(q, 0.85)
        for r, (f, rs, ps, mm) in enumerate(gathered):
            ref, ref_p = whole.register_frames(f, pose=pose)
            same = ref.tobytes() == rs.tobytes() and ref_p.tobytes() == ps.tobytes() and mm.tobytes() == m_ref.tobytes()
            print(f"rank {r}: sharded == unsharded: {same}; found {rs['found'].tolist()} poses {ps['found'].tolist()} matches {len(mm)}")
            ok = ok and same and rs["found"][0] == 1 and ps["found"][0] == 1
        whole.close()
        # kernel attributes are per device — a second context on another GPU of the same process must work
        c1 = dunk.Context(1, 2)
        a = np.random.default_rng(1).integers(0, 256, (300, 61), dtype=np.uint8)
        b = np.random.default_rng(2).integers(0, 256, (4000, 61), dtype=np.uint8)
        i0, d0 = dunk.feature_extraction.knn2(a, b, ctx)
        i1, d1 = dunk.feature_extraction.knn2(a, b, c1)
        e0 = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(tiles[3], None, ctx)
        e1 = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(tiles[3], None, c1)
        knn_same = np.array_equal(i0, i1) and np.array_equal(d0, d1)
        kp_same = e0.keypoints.tobytes() == e1.keypoints.tobytes()
        desc_same = e0.descriptors.tobytes() == e1.descriptors.tobytes()
        two = knn_same and kp_same and desc_same and len(e0.keypoints) > 100
        print(f"contexts on two devices in one process agree: {two} (knn {knn_same}, keypoints {kp_same}, descriptors {desc_same})")
        ok = ok and two
        c1.close()
    flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", local))
    dist.broadcast(flag, 0)
    torch.cuda.synchronize()
    shard.close(); built.close(); geo.close(); group.close(); ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAILED")
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
