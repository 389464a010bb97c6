"""One rank of the 2-GPU hardware test (tests/test_gpu_distributed.py): every multi-GPU plan of distributed.py on
this rank's shard, results written for the parent to compare with the oracle.  Launched by torch.distributed.run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_BUILD, N_PROBE, GROUPS = 40_000, 300_000, 997


def tables(start_l, n_l, start_r, n_r):
    from oracle import oracle as O
    lk = O.gen_perm_i64(start_l, n_l, 7368787 % N_BUILD or 1, N_BUILD)
    fk = O.gen_mod_i64(47, start_r, n_r, int(N_BUILD * 1.25))  # some probe rows find no match
    return lk, lk % GROUPS, fk, O.gen_unif_f64(48, start_r, n_r, 100.0)


def main():
    out_dir = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from importlib import import_module
    import nqe_b200 as nq
    D = import_module("naive-query-engine_b200.distributed")
    ctx = nq.Context(int(os.environ["LOCAL_RANK"]))  # NOT bound to torch's stream by the caller: the engine binds itself
    engine = D.CudaEngine(nq, ctx, torch)
    nl, nr = N_BUILD // world, N_PROBE // world
    lk, la, fk, rb = tables(rank * nl, nl, rank * nr, nr)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy()).cuda()
    lcols, rcols = [dev(lk), dev(la)], [dev(fk), dev(rb)]
    res = {}
    ph = D.Phases(torch)
    merged, wire = D.broadcast_join_group_by(dist, torch, engine, lcols, rcols, world, phases=ph)
    res["broadcast"], res["broadcast_wire"] = torch.stack(merged).cpu().numpy(), wire
    res["broadcast_phases"] = np.array(sorted(ph.ms()), dtype="U32")
    merged, wire = D.shuffled_join_group_by(dist, torch, engine, lcols, rcols, world, None)
    res["nccl"], res["nccl_wire"] = torch.stack(merged).cpu().numpy(), wire
    xbufs = (engine.alloc_exchange(N_BUILD, 2, dist.group.WORLD), engine.alloc_exchange(N_PROBE, 2, dist.group.WORLD))
    for _ in range(3):  # repeated: the receive buffers are reused, the barriers must order the exchanges
        merged, wire = D.shuffled_join_group_by(dist, torch, engine, lcols, rcols, world, xbufs)
    res["peer"], res["peer_wire"] = torch.stack(merged).cpu().numpy(), wire
    gk = dev((fk % 301).astype(np.int64))
    full, sent = D.distributed_group_by(dist, torch, engine, [gk, rcols[1]], world, gather=True)
    res["group_by"], res["group_by_sent"] = torch.stack(full).cpu().numpy(), sent
    jb, _ = D.distributed_hash_join(dist, torch, engine, lcols, rcols, world, plan="broadcast")
    js, _ = D.distributed_hash_join(dist, torch, engine, lcols, rcols, world, plan="shuffle")
    jp, _ = D.distributed_hash_join(dist, torch, engine, lcols, rcols, world, plan="shuffle", xbufs=xbufs)
    res["join_broadcast"] = torch.stack(jb).cpu().numpy() if jb[0].numel() else np.zeros((4, 0), np.int64)
    res["join_shuffle"] = torch.stack(js).cpu().numpy() if js[0].numel() else np.zeros((4, 0), np.int64)
    res["join_shuffle_peer"] = torch.stack(jp).cpu().numpy() if jp[0].numel() else np.zeros((4, 0), np.int64)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
