"""-m gpu, needs >= 2 GPUs (run with `gpurun --gpus 2`; skipped on a 1-GPU box): the ray-sharded render
over real ranks -- `dist.render_sharded` through the fused pixel exchange (kernel epilogue stores into
the peers' symmetric-memory buffers + lrf_peer_barrier) and through the NCCL all-gather fallback --
must equal the unsharded render bit for bit."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev,
                            timeout=datetime.timedelta(seconds=90))       # a rank that dies must not stall its peer for minutes
    res = {}
    try:
        import bench
        from localrf_b200.dist import PixelExchange, render_sharded
        lt = bench.build_scene(dev, 96)
        view = torch.tensor([0], device=dev)
        xch = PixelExchange(20000, device=dev)
        res["multicast"] = bool(xch.mc_ptr)
        ok_x, ok_n = True, True
        with torch.no_grad():
            for step, (lo, n) in enumerate([(0, 4096), (123456, 4096), (5000, 1001), (640000 - 17, 17),
                                            (300000, 20000), (1000, 5)]):
                ids = torch.arange(lo, lo + n, dtype=torch.int64, device=dev)
                rgb, depth, _, _ = lt(ids, view, 800, 800, is_train=False)
                r1, d1 = render_sharded(lt, ids, view, 800, 800, exchange=xch, is_train=False)
                ok_x &= bool(torch.equal(r1, rgb) and torch.equal(d1, depth))
                r2, d2 = render_sharded(lt, ids, view, 800, 800, is_train=False)
                ok_n &= bool(torch.equal(r2, rgb) and torch.equal(d2, depth))
        # lag 1 (no barrier kernel at all: the render kernel waits / publishes): step i returns step i-2
        xl = PixelExchange(8192, device=dev, lag=1)
        hist = []
        with torch.no_grad():
            for lo in (0, 4096, 77 * 4096, 100 * 4096, 8192, 50 * 4096, 3 * 4096):
                ids = torch.arange(lo, lo + 4096, dtype=torch.int64, device=dev)
                rgb, depth, _, _ = lt(ids, view, 800, 800, is_train=False)
                r1, d1 = render_sharded(lt, ids, view, 800, 800, exchange=xl, is_train=False)
                hist.append((rgb, depth))
                if len(hist) >= 3:
                    ok_x &= bool(torch.equal(r1, hist[-3][0]) and torch.equal(d1, hist[-3][1]))
        # unicast P2P stores as well when the multicast path was taken above
        if xch.mc_ptr:
            xu = PixelExchange(8192, device=dev, use_multicast=False)
            with torch.no_grad():
                for lo in (0, 77 * 4096):
                    ids = torch.arange(lo, lo + 4096, dtype=torch.int64, device=dev)
                    rgb, depth, _, _ = lt(ids, view, 800, 800, is_train=False)
                    r1, d1 = render_sharded(lt, ids, view, 800, 800, exchange=xu, is_train=False)
                    ok_x &= bool(torch.equal(r1, rgb) and torch.equal(d1, depth))
        torch.cuda.synchronize()
        res["exchange"], res["nccl"] = ok_x, ok_n
    except Exception as e:  # pragma: no cover
        import traceback
        res["error"] = f"{e}\n{traceback.format_exc()}"
    finally:
        q.put((rank, res))
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_render_sharded_equals_unsharded(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    print(results)
    for r in range(world):
        assert "error" not in results[r], results[r]["error"]
        assert results[r]["exchange"] and results[r]["nccl"], results
