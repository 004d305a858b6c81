"""world_size-2 check of the data-parallel HOST logic on CPU (gloo): one process per rank, weights broadcast
once from rank 0, each rank runs the GAN step on its own shard, ONE flat all-reduce(SUM) per optimiser with the
1/world average folded into the Adam kernel (replaces nn.DataParallel, processor_v2.py:167-172).
The kernels run on the CPU kernel-logic emulator of tests/emu (test infrastructure); the NCCL path on
the GPUs uses the same Processor code with backend "nccl"."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_lib, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        torch.set_num_threads(1)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from speech2affective_gestures_b200 import _C
        _C._inject_for_tests(emu_lib, strict=False)
        import test_step as ts
        from common import O, inject_eps
        from speech2affective_gestures_b200 import ops
        dev = torch.device("cpu")
        torch.manual_seed(1000 + rank)  # different initial weights per rank: the ctor broadcast must fix that
        pr, c = ts.make_processor("tiny", 40, 12, dev, batch_size=2)
        if rank == 1:  # de-synchronise on purpose, then re-run the one-time broadcast
            with torch.no_grad():
                pr.s2ag_generator.flat_params.add_(0.123)
                pr.s2ag_discriminator.flat_params.mul_(1.5)
        pr._init_distributed()
        assert pr.world == world and pr.rank == rank
        G, D = pr.s2ag_generator, pr.s2ag_discriminator
        p0 = {"G": G.flat_params.clone(), "D": D.flat_params.clone()}
        for name in ("G", "D"):  # identical after the broadcast
            both = [torch.empty_like(p0[name]) for _ in range(world)]
            dist.all_gather(both, p0[name])
            assert torch.equal(both[0], both[1]), name
        # capture the local (pre-reduce) gradients
        local = {}
        orig = pr._allreduce_grads

        def spy(net):
            local["G" if net is G else "D"] = net.flat_grads.clone()
            orig(net)
            local[("G" if net is G else "D") + "_sum"] = net.flat_grads.clone()
        pr._allreduce_grads = spy
        batch, eps_list, rand_idx = O.synthetic_batch(2, 40, 12, 36267, 500 + rank)  # rank-specific shard
        pr.injected_rand_idx = rand_idx
        for n in (G, D, pr.trimodal_generator):
            n.train()
        inject_eps(eps_list)
        pr.forward_pass_s2ag(batch[0], batch[1], batch[2], batch[3], batch[4], train=True)
        lr = {"G": pr.lr_s2ag_gen, "D": pr.lr_s2ag_dis}
        for name, net in (("G", G), ("D", D)):
            parts = [torch.empty_like(local[name]) for _ in range(world)]
            dist.all_gather(parts, local[name])
            assert not torch.equal(parts[0], parts[1]), "shards must differ"
            tot = parts[0] + parts[1]
            assert torch.allclose(local[name + "_sum"], tot, rtol=1e-6, atol=1e-9), name
            # first Adam step with the averaged gradient: p - lr * g / (|g| + eps)
            gavg = tot / world
            want = p0[name] - lr[name] * gavg / (gavg.abs() + 1e-8)
            assert torch.allclose(net.flat_params, want, rtol=1e-4, atol=1e-7), name
            both = [torch.empty_like(net.flat_params) for _ in range(world)]
            dist.all_gather(both, net.flat_params.detach().clone())
            assert torch.equal(both[0], both[1]), name + " diverged across ranks"
        # BatchNorm statistics stay per-rank (DataParallel semantics): different shards => different running means
        rm = G.state_dict()["aff_encoder.batch_norm1.running_mean"].clone()
        both = [torch.empty_like(rm) for _ in range(world)]
        dist.all_gather(both, rm)
        assert not torch.equal(both[0], both[1])
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


def test_two_rank_gan_step_gloo():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    emu_lib = build_emu.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_lib, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)
