"""Worker for the world_size>1 tests (launched with torch.distributed.run).

mode "gloo": CPU, host-side logic only — every rank produces its strip of a frame with the
  oracle, strips are gathered to rank 0 with the product's gather_strips_to_rank0 and compared
  with the whole-frame oracle render; pose blocks are checked to tile the batch.
mode "nccl": GPU — the same through the CUDA path (strip draws into torch-owned framebuffers,
  NCCL gather over NVLink), the strip group (rank 0's framebuffers shared over CUDA IPC, tiles pushed over
  NVLink, device-side hand-off flags), plus a frame-parallel batch compared per rank with the oracle.
mode "ipc1": the strip group with every rank on GPU 0 (CUDA IPC works between processes on one device; the
  process group is gloo, which only carries the handles and the barriers) — what a one-GPU box can run.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gorender_b200 as g  # noqa: E402
from gorender_b200 import parallel, workloads  # noqa: E402
from oracle_binding import Oracle  # noqa: E402
import scene_defs  # noqa: E402


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    orc = Oracle()
    sc = scene_defs.multi_object()
    H, W = sc.height, sc.width

    if mode == "gloo":
        dist.init_process_group("gloo")
        r = sc.renderer(None)
        ref = orc.draw(r, sc.objects, sc.camera)
        color = torch.zeros((H, W, 4), dtype=torch.uint8)
        depth = torch.zeros((H, W), dtype=torch.float32)
        y0, y1 = parallel.strip_rows(H, world, rank)
        color[y0:y1] = torch.from_numpy(ref["pixels"][y0:y1])
        depth[y0:y1] = torch.from_numpy(ref["zbuffer"][y0:y1])
        parallel.gather_strips_to_rank0(color, depth, H)
        if rank == 0:
            assert np.array_equal(color.numpy(), ref["pixels"])
            assert np.array_equal(depth.numpy().view(np.uint32), ref["zbuffer"].view(np.uint32))
        # frame-parallel partition: blocks tile the batch
        b, e = parallel.pose_block(37, world, rank)
        t = torch.tensor([b, e], dtype=torch.int64)
        out = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(out, t)
        blocks = [tuple(x.tolist()) for x in out]
        assert blocks == parallel.frame_parallel_blocks(37, world), blocks
        assert blocks[0][0] == 0 and blocks[-1][1] == 37
        # balanced strips: computed independently on every rank from the same weights, identical everywhere
        wts = np.array([0, 0, 3, 9, 9, 4, 0, 1, 0, 0, 0, 7], np.float64)
        mine = parallel.balanced_strip_rows(wts, world, 12 * 32 - 5)
        box = [None] * world
        dist.all_gather_object(box, mine)
        assert all(b == mine for b in box)
        dist.barrier()
        dist.destroy_process_group()
        print(f"rank {rank} ok")
        return

    one_gpu = mode == "ipc1"
    if one_gpu:
        local = 0
    torch.cuda.set_device(local)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    dev = g.Device(local, stream.cuda_stream)
    with torch.cuda.stream(stream):
        if not one_gpu:
            # ---- sort-first strips, gathered to rank 0 over NCCL
            tfb = parallel.TorchFrameBuffer(W, H, 1, dev, torch.device("cuda", local))
            tfb.color.zero_()
            tfb.depth.zero_()
            r = sc.renderer(tfb.fb)
            packed = r.pack_objects(sc.objects, [sc.camera])
            parallel.draw_strip(r, packed, H, world, rank)
            parallel.gather_strips_to_rank0(tfb.color[0], tfb.depth[0], H)
            stream.synchronize()
            if rank == 0:
                ref = orc.draw(r, sc.objects, sc.camera)
                assert np.array_equal(tfb.color[0].cpu().numpy(), ref["pixels"]), "gathered colour differs"
                assert np.array_equal(tfb.depth[0].cpu().numpy().view(np.uint32), ref["zbuffer"].view(np.uint32))
        # ---- sort-first strips the B200 way: rank 0's framebuffers shared over CUDA IPC, every rank's raster kernel
        #      writes its rows into them over NVLink, device-side flags hand the frame over; strips balanced by the
        #      busy tiles of a probe frame; double-buffered; rank 0 mirrors each frame into host memory
        probe = g.FrameBuffer(W, H, 1, dev)
        pr = sc.renderer(probe)
        packed = pr.pack_objects(sc.objects, [sc.camera])
        pr.draw_packed(packed, 0)
        weights = probe.tile_flags(0).sum(axis=1)
        rows = parallel.balanced_strip_rows(weights, world, H)
        assert rows[0][0] == 0 and rows[-1][1] == H and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        grp = parallel.StripGroup(dev, W, H, nbuf=2, rows=rows)
        for rr_ in grp.renderers:
            for k_, v_ in sc.options.items():
                setattr(rr_, k_, v_)
        base = [o.Translation.copy() for o in sc.objects]
        frames = []
        for f in range(5):
            for o, b0 in zip(sc.objects, base):
                o.Translation = (b0 + np.array([0.15 * f, -0.1 * f, 0], np.float32)).astype(np.float32)
            frames.append(np.ascontiguousarray(grp.renderers[0].pack_objects(sc.objects, [sc.camera])))
        got = []
        for f in range(5):
            k = f & 1
            grp.draw(k, frames[f])
            tpf = torch.tensor([int(0)], dtype=torch.int64, device="cpu" if one_gpu else "cuda")
            if rank == 0:
                fbk = grp.fbs[k]
                fbk.update_mirrors_async(0, 1, fbk.mirror("Pixels"), fbk.mirror("ZBuffer"))   # reads the whole frame, after every rank's flag
                grp.release(k)
                fbk.mirror("Pixels").wait()
                fbk.mirror("ZBuffer").wait()
                got.append((fbk.Pixels.copy(), fbk.ZBuffer.copy()))
            st = np.zeros(1, dtype=g._cabi.STATS_DTYPE)
            dev.check(dev.lib.grb_frame_stats_read(dev.h, 1, st.ctypes.data))
            y0, y1 = rows[rank]
            tpf[0] = int(st["tpf"][0]) if y1 > y0 else 0
            stream.synchronize()
            dist.all_reduce(tpf)
            if rank == 0:
                for o, b0 in zip(sc.objects, base):
                    o.Translation = (b0 + np.array([0.15 * f, -0.1 * f, 0], np.float32)).astype(np.float32)
                ref = orc.draw(grp.renderers[0], sc.objects, sc.camera)
                assert np.array_equal(got[f][0], ref["pixels"]), f"strip group frame {f}: colour differs"
                assert np.array_equal(got[f][1].view(np.uint32), ref["zbuffer"].view(np.uint32)), f"strip group frame {f}: depth differs"
                assert int(tpf[0]) == ref["tpf"], (int(tpf[0]), ref["tpf"])     # the ranks' TPFs add up to the frame's
        assert dev.signal_timeouts() == 0
        for o, b0 in zip(sc.objects, base):
            o.Translation = b0
        grp.close()
        probe.close()
        if one_gpu:
            dist.barrier()
            dist.destroy_process_group()
            print(f"rank {rank} ok")
            return
        # ---- frame-parallel: each rank renders its block of poses
        objs, cams = workloads.config_c5(n=16, poses=10)
        b, e = parallel.pose_block(len(cams), world, rank)
        fb = g.FrameBuffer(640, 360, max(e - b, 1), dev)
        rr = g.Renderer(fb)
        if e > b:
            px, z, tpf = rr.DrawBatch(objs, cams[b:e])
            for k in (0, e - b - 1):
                ref = orc.draw(rr, objs, cams[b + k])
                assert int(tpf[k]) == ref["tpf"]
                assert np.array_equal(px[k], ref["pixels"]) and np.array_equal(z[k].view(np.uint32), ref["zbuffer"].view(np.uint32))
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
