"""
NVLink peer-memory bandwidth of the exchange-copy kernel: push (local -> peer stores) versus
pull (peer loads -> local), run with torchrun on >= 2 GPUs.  Rank 0 prints one line per size.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-pme_b200"))
from torchpme_b200 import _native  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    n_bytes = 64 << 20
    buf = _native.PeerBuffer(n_bytes, dev)
    mine = torch.tensor(list(buf.handle), dtype=torch.uint8, device=dev)
    handles = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(handles, mine)
    peer = (rank + 1) % world
    peer_ptr = buf.open_peer(bytes(handles[peer].cpu().tolist()))
    local = torch.empty(n_bytes // 4, dtype=torch.float32, device=dev)
    dist.barrier()
    for mb in (1, 4, 8, 16, 32, 64):
        words = (mb << 20) // 8      # complex-float elements
        for chunks in (64, 256, 1024):
            run = words // chunks
            res = {}
            for mode in ("push", "pull", "local"):
                if mode == "push":
                    src, dst = local, peer_ptr
                elif mode == "pull":
                    src, dst = buf.as_tensor(0, (n_bytes // 4,), torch.float32), local.data_ptr()
                    # read the PEER's buffer: build a tensor view on the mapped pointer
                    class R: pass
                    r = R(); r.__cuda_array_interface__ = {"shape": (n_bytes // 4,), "typestr": "<f4", "data": (peer_ptr, False), "version": 2}
                    src = torch.as_tensor(r, device=dev)
                else:
                    src, dst = local, buf.ptr
                def go():
                    _native.slab_exchange_copy(src, [dst], 1, 1, chunks, run, (0, 0, run), (0, run))
                for _ in range(3):
                    go()
                torch.cuda.synchronize(); dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    go()
                b.record(); torch.cuda.synchronize()
                res[mode] = (mb << 20) * 10 / (a.elapsed_time(b) * 1e-3) / 1e9
                dist.barrier()
            if rank == 0:
                print(f"{mb:3d} MiB in {chunks:5d} chunks: push {res['push']:7.1f} GB/s  pull {res['pull']:7.1f} GB/s  "
                      f"local {res['local']:7.1f} GB/s (all ranks copy to/from their right neighbour at once)", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
