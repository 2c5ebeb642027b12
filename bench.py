#!/usr/bin/env python
"""bench.py -- SAGE-ResBN (configs/rec_k8.yaml) full-batch train step on the synthetic
Elliptic-shaped graph: epoch ms and GEdges/s fwd+bwd, with the SpMM roofline beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one reference `train_epoch` body
(src/train_gnn.py:187-209): forward over all nodes, masked weighted CE, backward, global-norm
clip, Adam.  N > 1 (launched by torch.distributed.run): weak scaling -- the graph is N
block-diagonal replicas of the Elliptic-shaped graph, timestep-sharded over the ranks with zero
halo; NCCL all-reduces only the flat weight-gradient buffer and the BatchNorm statistics.
`--impl reference`: the reference's CPU path (restated PyG nets, oracle/) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(arch="sage_resbn", hidden_dim=64, layers=3, dropout=0.20, weight_decay=5.0e-5, lr=5.0e-4,
           grad_clip=1.0, symmetrize_edges=True, time_embed_dim=2, time_embed_type="sin", max_timestep=49,
           train_window_k=8)  # /root/reference/configs/rec_k8.yaml
METRIC = "SAGE-ResBN full-batch train step throughput (fwd+bwd+clip+Adam), GEdges/s"
UNIT = "GEdges/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the roofline kernel, per launch, from the committed
    `ncu --set full` capture (profiles/); None when no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "r02", "roofline_traffic.json")
    try:
        d = json.load(open(p))
        return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])
    except Exception:
        return None


def ncu_l2_hit():
    """L2 hit rate (%) of the roofline kernel from the same committed capture; None when not recorded."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "r02", "roofline_traffic.json")))["l2_hit_rate_pct"])
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.max_mhz = index, period, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def host_graph(n_replicas: int):
    from egnn_b200 import synthetic
    gr = synthetic.make_elliptic_like(train_window_k=CFG["train_window_k"])
    if n_replicas > 1:
        gr = synthetic.replicate(gr, n_replicas)
    return gr


def spmm_bytes(n, f, e, in_es, out_es):
    """Compulsory (algorithmic) bytes of one mean-SpMM: read every source row once, write every
    output row once, plus col indices and row pointers."""
    return n * f * in_es + n * f * out_es + 4 * e + 4 * (n + 1)


# BASELINE.json configs other than the headline one (reference configs/*.yaml); timed at N=1 and reported in `configs`
OTHER_CONFIGS = {
    "gcn": dict(arch="gcn", hidden_dim=128, layers=3, dropout=0.5, lr=1e-3, weight_decay=5e-4, k=10, sym=False, amp=False,
                what="configs/gcn.yaml: GCN 167->128->128->2, self-loop graph E'=438124, fp32 (config 1 is the CPU run)"),
    "sage": dict(arch="sage", hidden_dim=128, layers=2, dropout=0.5, lr=1e-3, weight_decay=5e-4, k=10, sym=True, amp=False,
                 what="configs/sage.yaml + symmetrize_edges: SAGE 167->128->2, fp32"),
    "gat": dict(arch="gat", hidden_dim=32, layers=2, heads=4, dropout=0.5, lr=1e-3, weight_decay=5e-4, k=10, sym=False,
                amp=True, what="configs/gat.yaml: GAT 167->(4x8)->2, self-loop graph, bf16 autocast (attention fp32)"),
    "sage_l3": dict(arch="sage", hidden_dim=128, layers=3, dropout=0.4, lr=1e-3, weight_decay=5e-4, k=18, sym=True,
                    amp=True, what="configs/sage_l3_k18.yaml: SAGE 167->128->128->2 on the base graph (x1), bf16 autocast"),
}


def _timed(fn, n, barrier=None):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def run_other_configs(dev, steps, peak):
    """Step time of BASELINE configs 1, 2, 4 and 5 (x1) at N=1: CUDA-graph train step on the same synthetic graph, plus
    the config's own dominant aggregation launch against the HBM roofline and, for gcn.yaml (defined as the CPU run),
    the oracle's CPU step beside it."""
    import egnn_b200 as E
    from egnn_b200 import _lib, ops, synthetic
    from egnn_b200.train import TrainStep
    out = []
    for name, cfg in OTHER_CONFIGS.items():
        gr = synthetic.make_elliptic_like(train_window_k=cfg["k"])
        x = torch.cat([gr.x, (gr.timestep.float() / gr.timestep.max().float()).unsqueeze(1)], dim=1)  # use_time_scalar
        ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1) if cfg["sym"] else gr.edge_index
        torch.manual_seed(42)
        model = E.build_model(cfg["arch"], x.size(1), cfg).to(dev)
        model.set_dropout_seed(42, dev)
        step = TrainStep(model, x.to(dev), ei.to(dev), gr.timestep.to(dev), gr.y.to(dev), gr.train_mask.to(dev),
                         lr=cfg["lr"], weight_decay=cfg["weight_decay"], grad_clip=1.0, amp=cfg["amp"])
        step.run()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        step.run()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - n0
        step.capture(warmup=2)
        for _ in range(3):
            step.run()
        ms = _timed(step.run, steps)
        self_loops = cfg["arch"] in ("gcn", "gat")
        e_used = ei.size(1)
        ent = {"name": name, "workload": cfg["what"], "dtype": "bf16" if cfg["amp"] else "f32",
               "ms_per_step": round(ms, 4), "gedges_per_s": round(e_used / (ms * 1e-3) / 1e9, 4), "edges": int(e_used),
               "gpu_launches_per_step": int(launches), "loss": round(float(step.loss), 6)}
        # the config's widest aggregation launch, alone, rotating inputs (algorithmic bytes as in SURVEY 8d)
        g = E.cached_graph(step.edge_index, gr.num_nodes, self_loops=self_loops)
        N = gr.num_nodes
        if cfg["arch"] == "gcn":
            F, idt, odt, mode, es_in, es_out, extra = 128, torch.float32, torch.float32, _lib.SPMM_WEIGHTED, 4, 4, 4
            e_k, label = N + gr.edge_index.size(1), "weighted (gcn_norm) F=128 fp32"
        elif cfg["arch"] == "gat":
            F = None
        elif name == "sage":
            F, idt, odt, mode, es_in, es_out, extra = 168, torch.float32, torch.float32, _lib.SPMM_MEAN, 4, 4, 0
            e_k, label = e_used, "mean F=168 (167 padded) fp32"
        else:
            F, idt, odt, mode, es_in, es_out, extra = 128, torch.bfloat16, torch.bfloat16, _lib.SPMM_MEAN, 2, 2, 0
            e_k, label = e_used, "mean F=128 bf16"
        if F is not None:
            n_rot = 6 if es_in == 4 else 8
            ins = [torch.randn(N, F, device=dev).to(idt) for _ in range(n_rot)]
            o = torch.empty(N, F, dtype=odt, device=dev)
            it = iter(range(10 ** 9))
            fn = lambda: ops.spmm(g, "csr", mode, ins[next(it) % n_rot], odt, out=o)
            for _ in range(n_rot):
                fn()
            t = _timed(fn, 20)
            b = spmm_bytes(N, F, e_k, es_in, es_out) + extra * e_k
            ent["spmm"] = {"kernel": label, "us": round(t * 1e3, 2), "algorithmic_bytes": int(b),
                           "frac": round(b / t / 1e6 / peak, 4)}
            del ins, o
        if name == "gcn":
            from oracle import pyg_restated as O
            torch.set_num_threads(os.cpu_count() or 1)
            torch.manual_seed(42)
            ref = O.build_model(cfg["arch"], x.size(1), cfg)
            opt = torch.optim.Adam(ref.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
            cw = O.class_weight(gr.y[gr.train_mask])
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                O.train_step(ref, x, ei, gr.timestep, gr.y, gr.train_mask, cw, opt, 1.0)
                ts.append(time.perf_counter() - t0)
            ent["cpu_ms_per_step"] = round(min(ts) * 1e3, 1)
            ent["cpu_cores"] = torch.get_num_threads()
        out.append(ent)
        del step, model, x, ei, g
        E.graph._GLOBAL_CACHE.clear()
        torch.cuda.empty_cache()
    return out


def run_minibatch(dev, with_cpu: bool):
    """The mini-batch path (`mini_batch: true`, src/train_gnn.py:329-348,212-245) with the reference's defaults -- fanout
    [10, 10], batch_size 8192 -- on the rec_k8 graph: per-batch time of the device-side NeighborLoader (sampling +
    relabelling + row slices, one synchronisation per batch), batch 0 compared bit for bit with the sequential CPU
    restatement of PyG's sampler (timed beside it), and the wall time of one mini-batch training epoch."""
    import time as _time
    import egnn_b200 as E
    from egnn_b200 import _lib, synthetic
    from egnn_b200.train import class_weight, train_epoch_minibatch
    gr = synthetic.make_elliptic_like(train_window_k=CFG["train_window_k"])

    class _D:
        pass
    d = _D()
    d.x, d.y, d.timestep = gr.x.to(dev), gr.y.to(dev), gr.timestep.to(dev)
    d.train_mask = gr.train_mask.to(dev)
    ei_h = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    d.edge_index = ei_h.to(dev)
    idx = torch.nonzero(d.train_mask).view(-1)
    fan, bs = [10, 10], 8192
    mk = lambda: E.NeighborLoader(d, num_neighbors=fan, batch_size=bs, input_nodes=idx, shuffle=False, seed=42)
    loader = mk()
    first = next(iter(loader))
    for _ in loader:      # warm-up epoch
        pass
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    t0 = _time.perf_counter()
    nb = nn = ne = 0
    for _ in range(5):
        for b in loader:
            nb, nn, ne = nb + 1, nn + b.num_nodes, ne + int(b.edge_index.size(1))
    torch.cuda.synchronize()
    ms_batch = (_time.perf_counter() - t0) / nb * 1e3
    out = {"workload": f"NeighborLoader(num_neighbors={fan}, batch_size={bs}, input_nodes=train_idx) on the rec_k8 graph "
                       f"({idx.numel()} seed nodes, {len(loader)} batches per epoch)",
           "ms_per_batch": round(ms_batch, 4), "nodes_per_batch": nn // nb, "edges_per_batch": ne // nb,
           "gpu_launches_per_batch": int((_lib.launch_count() - n0) // nb),
           "includes": "multi-hop sampling, relabelling, x / y / timestep / mask row slices, one 8-byte read-back"}
    if with_cpu:
        from oracle.neighbor_sample_np import csc_by_destination, neighbor_sample
        ip, src, eid = csc_by_destination(ei_h.numpy(), gr.num_nodes)
        t0 = _time.perf_counter()
        n_id, le, e_id, _, _ = neighbor_sample(ip, src, eid, idx[:bs].cpu().numpy(), fan, seed=42, batch_idx=0)
        cpu_ms = (_time.perf_counter() - t0) * 1e3
        import numpy as _np
        out["parity"] = {"what": "batch 0 vs the sequential CPU restatement of PyG's sampler (same Philox stream): "
                                 "node list, local edge list, edge ids",
                         "bit_exact": bool(_np.array_equal(first.n_id.cpu().numpy(), n_id)
                                           and _np.array_equal(first.edge_index.cpu().numpy(), le)
                                           and _np.array_equal(first.e_id.cpu().numpy(), e_id))}
        out["cpu_ms_per_batch"] = round(cpu_ms, 1)
        out["cpu_kind"] = "port (pure-Python sequential restatement, 1 core; sampling only, no row slices)"
    torch.manual_seed(0)
    model = E.build_model("sage_resbn", 166, {k: CFG[k] for k in ("hidden_dim", "layers", "dropout", "time_embed_dim",
                                                                   "time_embed_type", "max_timestep") if k in CFG}).to(dev)
    loss_fn = E.make_loss_fn({}, class_weight(d.y[d.train_mask]), model, 1, 49)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=5e-5)
    loader = E.NeighborLoader(d, num_neighbors=fan, batch_size=bs, input_nodes=idx, shuffle=True, seed=42)
    for _ in range(2):      # warm-up epochs (batch shapes change with the shuffle: the caching allocator settles)
        train_epoch_minibatch(model, loader, opt, loss_fn, {"grad_clip": 1.0}, use_amp=True)
    times = []
    for _ in range(7):
        torch.cuda.synchronize()
        t0 = _time.perf_counter()
        loss = train_epoch_minibatch(model, loader, opt, loss_fn, {"grad_clip": 1.0}, use_amp=True)
        torch.cuda.synchronize()
        times.append((_time.perf_counter() - t0) * 1e3)
    out["epoch_ms"] = round(sorted(times)[len(times) // 2], 3)      # median of 7 epochs (wall)
    out["epoch_ms_min_max"] = [round(min(times), 3), round(max(times), 3)]
    out["epoch_loss"] = round(float(loss), 6)
    out["epoch_what"] = "train_epoch_minibatch, rec_k8 net, bf16 autocast, eager (per batch: sample, graph build, fwd, bwd, clip, Adam)"
    E.graph._GLOBAL_CACHE.clear()
    return out


def run_sage_l3_x64(dev, rank, world, steps, barrier, replicas=64):
    """BASELINE config 5 / north_star 'replicated scale-up': 3-layer SAGE (configs/sage_l3_k18.yaml) on 64 block-diagonal
    replicas of the Elliptic-shaped graph (N = 13 041 216, E' = 29 997 440), STRONG scaling: the 64 x 49 (replica, timestep)
    blocks are split over the ranks as contiguous replica ranges (zero halo), the flat weight gradient is all-reduced.
    Replica features are drawn on the device from per-replica seeds, so the global graph is the same for every N."""
    import torch.distributed as dist
    import egnn_b200 as E
    from egnn_b200 import synthetic
    from egnn_b200.shard import Shard, ShardedContext
    from egnn_b200.train import TrainStep
    cfg = dict(arch="sage", hidden_dim=128, layers=3, dropout=0.4, lr=1e-3, weight_decay=5e-4)
    if replicas % world:
        return {"skipped": f"{replicas} replicas do not divide over {world} ranks"}
    base = synthetic.make_elliptic_like(train_window_k=18)
    n1, kk = base.num_nodes, replicas // world
    r0 = rank * kk
    x = torch.empty((kk * n1, 168), dtype=torch.float32, device=dev)     # 167 = 166 + scalar time, padded to 16-byte rows
    tcol = (base.timestep.float() / base.timestep.max().float()).to(dev)
    for j in range(kk):
        r = r0 + j
        blk = x[j * n1:(j + 1) * n1]
        if r == 0:
            blk[:, :166] = base.x.to(dev)
        else:
            gen = torch.Generator(device=dev).manual_seed(42 + r)
            blk[:, :166] = synthetic._features(n1, 166, gen)
        blk[:, 166] = tcol
        blk[:, 167] = 0.0
    x = x[:, :167]                                                          # a strided view: rows stay 16-byte aligned
    eb = base.edge_index.to(dev)
    ei = torch.cat([eb + j * n1 for j in range(kk)], dim=1)
    ei_sym = torch.cat([ei, ei.flip(0)], 1).contiguous()
    rep = lambda t: t.repeat(kk)
    local_g = synthetic.EllipticGraph(x=x[:0].cpu(), edge_index=ei[:, :0].cpu(), y=rep(base.y), timestep=rep(base.timestep),
                                      train_mask=rep(base.train_mask), val_mask=rep(base.val_mask),
                                      test_mask=rep(base.test_mask))
    sh = Shard(rank=rank, world=world, row0=r0 * n1, n_local=kk * n1, n_global=replicas * n1, graph=local_g)
    ctx = ShardedContext(sh, dev)
    torch.manual_seed(42)
    model = ctx.attach(E.build_model(cfg["arch"], 167, cfg).to(dev))
    model.set_dropout_seed(42, dev)
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    step = TrainStep(model, x, ei_sym, local_g.timestep.to(dev), local_g.y.to(dev), local_g.train_mask.to(dev),
                     lr=cfg["lr"], weight_decay=cfg["weight_decay"], grad_clip=1.0, amp=True, cw=ctx.class_weight,
                     n_train_total=ctx.n_train_total, grad_reducer=ctx.reduce_grads if world > 1 else None,
                     health_check=ctx.check)
    del ei, eb
    step.run()
    torch.cuda.synchronize()
    step.capture(warmup=2)
    for _ in range(2):
        step.run()
    ms = _timed(step.run, steps, barrier)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    loss = step.loss.detach().clone().double()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(loss)
    step.loss_value()
    e_total = 2 * base.edge_index.size(1) * replicas
    ms = float(t)
    res = {"workload": "sage_l3_k18 (3-layer SAGE 167->128->128->2, bf16 autocast) on the 64x replicated graph",
           "scaling": "strong", "n_gpus": world, "replicas": replicas, "nodes_total": replicas * n1,
           "edges_total": int(e_total), "ms_per_step": round(ms, 4), "gedges_per_s": round(e_total / (ms * 1e-3) / 1e9, 4),
           "steps": steps, "loss_sum_over_ranks": round(float(loss), 6),
           "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
           "collectives": "peer-memory all-reduce kernel" if ctx.p2p else ("nccl" if world > 1 else "none")}
    del step, model, x, ei_sym
    E.graph._GLOBAL_CACHE.clear()
    torch.cuda.empty_cache()
    return res


def step1_parity(init_state, masks, host_gr, ei_host, loss_bf16_gpu, gnorm_bf16_gpu, loss_fp32_gpu, gnorm_fp32_gpu):
    """the first train step of THIS run (GPU, from the initial weights, its own Philox dropout masks)
    against the CPU oracle on the same graph, same weights, same masks -- fp32 and bf16-autocast."""
    from oracle import pyg_restated as O
    res = {}
    cw = O.class_weight(host_gr.y[host_gr.train_mask])
    for tag, amp_dtype in (("fp32", None), ("bf16", torch.bfloat16)):
        ref = O.build_model(CFG["arch"], host_gr.x.size(1), CFG)
        ref.load_state_dict(init_state)
        ref.train()
        if amp_dtype is not None:
            with torch.autocast(device_type="cpu", dtype=amp_dtype):
                logits = ref(host_gr.x, ei_host, host_gr.timestep, dropout_masks=masks)
                loss = O.masked_weighted_ce(logits.float(), host_gr.y, host_gr.train_mask, cw)
        else:
            logits = ref(host_gr.x, ei_host, host_gr.timestep, dropout_masks=masks)
            loss = O.masked_weighted_ce(logits, host_gr.y, host_gr.train_mask, cw)
        loss.backward()
        gn = float(torch.nn.utils.clip_grad_norm_(ref.parameters(), CFG["grad_clip"]))
        res[f"cpu_{tag}"] = {"loss": float(loss.detach()), "grad_norm": gn}
    rel = lambda a, b: abs(a - b) / abs(b)
    c32, c16 = res["cpu_fp32"], res["cpu_bf16"]
    return {"what": "step 1 from the initial weights, same graph / weights / dropout masks: GPU vs CPU oracle",
            "gpu_fp32_loss": loss_fp32_gpu, "gpu_bf16_loss": loss_bf16_gpu, "cpu_fp32_loss": c32["loss"],
            "cpu_bf16_loss": c16["loss"], "loss_rel_fp32": rel(loss_fp32_gpu, c32["loss"]),
            "loss_rel_bf16_vs_cpu_bf16": rel(loss_bf16_gpu, c16["loss"]),
            "loss_rel_bf16_vs_cpu_fp32": rel(loss_bf16_gpu, c32["loss"]),
            "cpu_bf16_vs_cpu_fp32": rel(c16["loss"], c32["loss"]),
            "grad_norm_rel_fp32": rel(gnorm_fp32_gpu, c32["grad_norm"]),
            "grad_norm_rel_bf16_vs_cpu_fp32": rel(gnorm_bf16_gpu, c32["grad_norm"]),
            "cpu_bf16_grad_norm_vs_cpu_fp32": rel(c16["grad_norm"], c32["grad_norm"]),
            "tolerance": {"fp32": 1e-5, "bf16": 4e-2},
            "ok": bool(rel(loss_fp32_gpu, c32["loss"]) <= 1e-5 and rel(loss_bf16_gpu, c32["loss"]) <= 4e-2
                       and rel(gnorm_fp32_gpu, c32["grad_norm"]) <= 1e-4)}


# ------------------------------------------------------------------------------- ours -------
def run_ours(args):
    import torch.distributed as dist
    import egnn_b200 as E
    from egnn_b200 import _lib, ops
    from egnn_b200.shard import ShardedContext, make_shard
    from egnn_b200.train import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stdout_fd = None
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; the contract is ONE JSON line,
        # so stdout is parked on stderr until the result line is written
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    amp = True
    gr = host_graph(world)
    sh = make_shard(gr, rank, world)
    lg = sh.graph
    ei_host = torch.cat([lg.edge_index, lg.edge_index.flip(0)], dim=1).contiguous()  # symmetrize_edges
    e_local, e_total = ei_host.size(1), 2 * gr.edge_index.size(1)
    del gr
    # pinned host buffers: the step's inputs as the reference holds them before data.to(device)
    host = {"x": lg.x, "ei": ei_host, "t": lg.timestep, "y": lg.y, "m": lg.train_mask}
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    devb = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()
    ctx = ShardedContext(sh, dev)
    torch.manual_seed(42)
    model = ctx.attach(E.build_model(CFG["arch"], lg.x.size(1), CFG).to(dev))
    model.set_dropout_seed(42, dev)
    if world > 1:  # identical initial weights on every rank
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    step = TrainStep(model, devb["x"], devb["ei"], devb["t"], devb["y"], devb["m"], lr=CFG["lr"],
                     weight_decay=CFG["weight_decay"], grad_clip=CFG["grad_clip"], amp=amp,
                     cw=ctx.class_weight, n_train_total=ctx.n_train_total,
                     grad_reducer=ctx.reduce_grads if world > 1 else None, health_check=ctx.check)
    parity_in = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        # step 1 of a throw-away fp32 twin and of the timed bf16 model, both from the same initial weights and the
        # same dropout stream: losses / gradient norms for the GPU-vs-CPU parity block of the result line
        import copy
        init_state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        twin = copy.deepcopy(model)
        twin.set_dropout_seed(42, dev)
        st32 = TrainStep(twin, devb["x"], devb["ei"], devb["t"], devb["y"], devb["m"], lr=CFG["lr"],
                         weight_decay=CFG["weight_decay"], grad_clip=CFG["grad_clip"], amp=False, cw=ctx.class_weight)
        l32 = float(st32.run())
        gn32 = float(st32.opt.grad_norm)
        del st32, twin
    step.run()                      # first call builds and caches the CSR/CSC structure
    torch.cuda.synchronize()
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        masks = [ops.dropout_mask(lg.num_nodes, CFG["hidden_dim"], CFG["dropout"], 42, li, seed_off=model._drop.offset).cpu()
                 for li in range(CFG["layers"] - 1)]
        parity_in = (init_state, masks, float(step.loss), float(step.opt.grad_norm), l32, gn32)
    n0 = _lib.launch_count()
    step.run()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - n0   # kernels of libegnn_b200.so per train step
    graphed = not args.eager
    try:
        if graphed:
            step.capture(warmup=max(1, args.warmup - 1))
        else:
            for _ in range(args.warmup):
                step.run()
    except Exception as ex:  # e.g. a collective that cannot be captured: stay eager, say so
        graphed = False
        step.graph = None
        if rank == 0:
            print(f"[bench] CUDA-graph capture failed ({type(ex).__name__}: {ex}); timing eager", file=sys.stderr)
        for _ in range(args.warmup):
            step.run()
    for _ in range(3):
        step.run()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    # ---- device-resident timed region: exactly K steps -------------------------------------
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step.run()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    ctx.check()       # a peer-memory all-reduce that timed out inside the timed region fails the run loudly
    # the same step WITHOUT the memoised layer-0 input layout ([x | sin | cos] fp32 + bf16 is re-derived from x and the
    # timesteps inside every step, as the reference's `_inject_time` does): reported beside the headline number
    ms_dyn = None
    if graphed:
        step.capture_dynamic()
        for _ in range(3):
            step.run(dynamic=True)
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step.run(dynamic=True)
        ev1.record()
        barrier()
        ms_dyn = ev0.elapsed_time(ev1) / args.steps
    # ---- end-to-end: host buffers in, loss out, every step ------------------------------------
    g_static = E.cached_graph(devb["ei"], lg.num_nodes)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    # Public API: egnn_b200.train.HostFeed -- every step copies its inputs from pinned host memory (staged on a copy
    # stream, so the copy of step i+1 overlaps the compute of step i), rebuilds CSR/CSC + row partition from the
    # new edge list, runs the step and reads the loss back (the read of step i completes while step i+1 runs).
    from egnn_b200.train import HostFeed
    feed = HostFeed(step, host, devb, lg.num_nodes, g_static)

    def e2e_run(n):
        feed.submit()
        if n > 1:
            feed.submit()       # one submission ahead: the copy of step i+2 is queued before run(i+1) wakes the host
        losses = []
        for i in range(n):
            prev = feed.run()
            if i + 2 < n:
                feed.submit()
            if prev is not None:
                losses.append(prev)
        losses.append(feed.drain())
        return losses

    e2e_run(3)
    e2e_steps = max(3, min(args.steps, 30))
    barrier()
    ev0.record()
    e2e_losses = e2e_run(e2e_steps)
    ev1.record()
    barrier()
    assert len(e2e_losses) == e2e_steps
    ms_e2e = ev0.elapsed_time(ev1)
    ctx.check()
    clocks = sampler.stop()
    # ---- eval_split forward (src/train_gnn.py:248-257): fp32, never under autocast, BatchNorm on running stats;
    # the reference's epoch = train step + this forward (eval_fwd_ms, ref_epoch_ms)
    from egnn_b200.train import eval_probs
    for _ in range(2):
        eval_probs(model, devb["x"], devb["ei"], devb["t"])
    barrier()
    ev0.record()
    for _ in range(5):
        eval_probs(model, devb["x"], devb["ei"], devb["t"])
    ev1.record()
    barrier()
    ms_eval_eager = ev0.elapsed_time(ev1) / 5
    # the same forward replayed as one CUDA graph (train.EvalStep; what metrics.fit runs per epoch): device-bound
    from egnn_b200.train import EvalStep
    evs = EvalStep(model, devb["x"], devb["ei"], devb["t"]).capture()
    for _ in range(3):
        evs.run()
    barrier()
    ev0.record()
    for _ in range(20):
        evs.run()
    ev1.record()
    barrier()
    ms_eval = ev0.elapsed_time(ev1) / 20
    del evs
    # ---- epoch tail on the device: validation PR-AUC from the eval logits + early-stopping
    # bookkeeping with the best-parameter snapshot (the reference does these on the host, src/train_gnn.py:387-402)
    from egnn_b200 import metrics as dev_metrics
    _, ev_logits = eval_probs(model, devb["x"], devb["ei"], devb["t"])
    ev_logits = ev_logits.float().contiguous()
    val_mask = lg.val_mask.to(dev)
    stopper = dev_metrics.EarlyStopper(patience=20, flat_param=step.opt.flat_param)
    ap_out = torch.empty(8, dtype=torch.float64, device=dev)
    for _ in range(2):
        dev_metrics.average_precision(devb["y"], val_mask, logits=ev_logits, out=ap_out)
        stopper.update(ap_out)
    barrier()
    ev0.record()
    for _ in range(5):
        dev_metrics.average_precision(devb["y"], val_mask, logits=ev_logits, out=ap_out)
        stopper.update(ap_out)
    ev1.record()
    barrier()
    ms_tail = ev0.elapsed_time(ev1) / 5
    val_ap = [float(v) for v in ap_out.tolist()]
    model.train()
    final_loss = float(step.loss)
    t = torch.tensor([ms, ms_e2e, ms_eval, ms_tail, ms_dyn or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_eval, ms_tail, ms_dyn = (float(v) for v in t.tolist())
    ms_step, ms_e2e_step = ms / args.steps, ms_e2e / e2e_steps

    # ---- roofline of the dominant sparse kernel (layer-0 mean SpMM, F=168, fp32 -> bf16), timed alone
    peak, peak_src = peaks()
    roof, kernels = None, []
    if rank == 0:
        g = g_static
        N = lg.num_nodes

        def time_kernel(fn, bufs, iters=20):
            for b in bufs:
                fn(b)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(iters):
                fn(bufs[i % len(bufs)])
            b_.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b_) / iters

        # rotate over 4 input copies (4 x 137 MB > 126 MB L2) so launches do not hit a warm L2
        xs168 = [torch.randn(N, 168, device=dev) for _ in range(4)]
        out168 = torch.empty(N, 168, dtype=torch.bfloat16, device=dev)
        t168 = time_kernel(lambda b: ops.spmm(g, "csr", _lib.SPMM_MEAN, b, torch.bfloat16, out=out168), xs168)
        b168 = spmm_bytes(N, 168, e_local, 4, 2)
        del xs168
        xs64 = [torch.randn(N, 64, device=dev).bfloat16() for _ in range(8)]
        out64 = torch.empty(N, 64, dtype=torch.bfloat16, device=dev)
        t64 = time_kernel(lambda b: ops.spmm(g, "csr", _lib.SPMM_MEAN, b, torch.bfloat16, out=out64), xs64)
        add64 = torch.randn(N, 64, device=dev).bfloat16()
        t64b = time_kernel(lambda b: ops.spmm(g, "csc", _lib.SPMM_SUM, b, torch.bfloat16, out=out64, addend=add64), xs64)
        b64 = spmm_bytes(N, 64, e_local, 2, 2)
        roof = {"kernel": "egnn_spmm mean fp32->bf16 F=168 (layer-0 aggregation, spmm_stream)", "bound": "hbm",
                "achieved": round(b168 / t168 / 1e6, 1), "peak": peak, "unit": "GB/s",
                "frac": round(b168 / t168 / 1e6 / peak, 4), "traffic": ncu_traffic(), "peak_source": peak_src,
                "algorithmic_bytes": b168, "us": round(t168 * 1e3, 2),
                # the gather-counted figure (every edge's source row counted, not each row once) and the
                # measured L2 hit rate beside the compulsory-traffic fraction -- the base graph is L2-scale
                "gather_counted_bytes": int(e_local * 168 * 4 + N * 168 * 2 + 4 * e_local + 4 * (N + 1)),
                "gather_counted_GBps": round((e_local * 168 * 4 + N * 168 * 2 + 4 * e_local + 4 * (N + 1)) / t168 / 1e6, 1),
                "l2_hit_rate_pct": ncu_l2_hit()}
        kernels = [
            {"kernel": "spmm mean fwd F=64 bf16", "us": round(t64 * 1e3, 2), "GBps": round(b64 / t64 / 1e6, 1)},
            {"kernel": "spmm transposed sum + addend (CSC) F=64 bf16", "us": round(t64b * 1e3, 2),
             "GBps": round(b64 / t64b / 1e6, 1)},
        ]

    # ---- the other BASELINE configs (N=1) and the 64x strong-scaling workload (every N) --------------------------
    rebuilds = feed.rebuilds
    del feed, step, model, host, devb, g_static
    E.graph._GLOBAL_CACHE.clear()
    torch.cuda.empty_cache()
    if world == 1 and rank == 0 and not args.skip_configs:
        # the same aggregation kernel on the 8x replicated graph: working set (1.1 GB of features at F=168) far beyond
        # the 126 MB L2, so the fraction is an HBM number whatever the cache does (the base graph is L2-scale)
        from egnn_b200 import synthetic
        g8h = synthetic.replicate(synthetic.make_elliptic_like(train_window_k=CFG["train_window_k"]), 8)
        ei8 = torch.cat([g8h.edge_index, g8h.edge_index.flip(0)], 1).to(dev)
        N8, e8 = g8h.num_nodes, ei8.size(1)
        g8 = E.build_graph(ei8, N8)
        for F, di, es_in in ((168, torch.float32, 4), (128, torch.bfloat16, 2), (64, torch.bfloat16, 2)):
            x8 = torch.randn(N8, F, device=dev).to(di)
            o8 = torch.empty(N8, F, dtype=torch.bfloat16, device=dev)
            fn = lambda: ops.spmm(g8, "csr", _lib.SPMM_MEAN, x8, torch.bfloat16, out=o8)
            for _ in range(3):
                fn()
            t8 = _timed(fn, 10)
            b8 = spmm_bytes(N8, F, e8, es_in, 2)
            kernels.append({"kernel": f"spmm mean fwd F={F} {'fp32' if es_in == 4 else 'bf16'}->bf16 on the 8x graph "
                                      f"(N={N8}, E'={e8})", "us": round(t8 * 1e3, 1), "algorithmic_bytes": b8,
                            "GBps": round(b8 / t8 / 1e6, 1), "frac": round(b8 / t8 / 1e6 / peak, 4)})
            del x8, o8
        del g8, ei8, g8h
        torch.cuda.empty_cache()
    other = None
    if world == 1 and rank == 0 and not args.skip_configs:
        other = run_other_configs(dev, max(10, min(args.steps, 30)), peak)
    minib = None
    if world == 1 and rank == 0 and not args.skip_configs:
        try:
            minib = run_minibatch(dev, with_cpu=not args.no_cpu_baseline)
        except Exception as ex:   # never lose the headline line to a secondary workload
            minib = {"failed": f"{type(ex).__name__}: {ex}"}
        torch.cuda.empty_cache()
    x64 = None
    if not args.skip_x64:
        try:
            x64 = run_sage_l3_x64(dev, rank, world, max(5, min(args.steps, 10)), barrier)
        except Exception as ex:   # never lose the headline line to the secondary workload
            x64 = {"failed": f"{type(ex).__name__}: {ex}"}
    if rank == 0:
        cpu = cpu_baseline_sample(budget_s=20.0) if (world == 1 and not args.no_cpu_baseline) else None
        parity = None
        if parity_in is not None:
            gr_h = host_graph(1)
            parity = step1_parity(parity_in[0], parity_in[1], gr_h,
                                  torch.cat([gr_h.edge_index, gr_h.edge_index.flip(0)], dim=1), *parity_in[2:])
        line = {
            "metric": METRIC, "value": round(e_total / (ms_step * 1e-3) / 1e9, 4), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
            "epoch_ms": round(ms_step, 4), "ms_per_step_inputs_rederived": round(ms_dyn, 4) if ms_dyn else None,
            "eval_fwd_ms": round(ms_eval, 4), "eval_fwd_eager_ms": round(ms_eval_eager, 4),
            "ref_epoch_ms": round(ms_step + ms_eval, 4), "epoch_tail_ms": round(ms_tail, 4),
            "val_pr_auc": {"value": round(val_ap[0], 6), "rows": int(val_ap[1]), "positives": int(val_ap[2]),
                           "where": "device (egnn_average_precision + egnn_early_stop_update)"},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "rec_k8: SAGE-ResBN 168->64->64->2, BN, residual, sin-2 time embed, "
                                   "symmetrize_edges, dropout 0.2, bf16 autocast, full-batch",
                       "nodes_per_gpu": lg.num_nodes, "edges_total": e_total, "replicas": world,
                       "parallelism": f"timestep-sharded dp{world}", "cuda_graph": graphed,
                       "collectives": ("peer-memory all-reduce kernel (csrc/p2p.cu) for BatchNorm statistics and "
                                       "gradients" if ctx.p2p else ("nccl" if world > 1 else "none")),
                       "l2": "inputs larger than L2 (x alone 135 MB; ~0.9 GB touched per step)",
                       "static_inputs": "x / timestep / edge_index are device-resident and unchanged between steps (the "
                                        "reference moves the graph to the device once): the sorted graph views and the "
                                        "layer-0 input layout [x | sin | cos] (fp32 + bf16) are memoised per tensor "
                                        "version; ms_per_step_inputs_rederived is the same step with the layout "
                                        "re-derived from x every step, and e2e (inputs re-copied every step) always "
                                        "re-derives it"},
            "clocks": clocks,
            "e2e": {"value": round(e_total / (ms_e2e_step * 1e-3) / 1e9, 4), "unit": UNIT,
                    "ms_per_step": round(ms_e2e_step, 4), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "steps": e2e_steps, "graph_rebuilds": rebuilds,
                    "includes": "every step: pinned-host x/edge_index/timestep/y/mask -> device (staged on a copy stream, "
                                "overlapping the previous step), byte-wise device comparison of the submitted edge_index "
                                "with the one the CSR/CSC views were built from (rebuild only when it differs), step, "
                                "loss -> host"},
            # edges visited by the aggregation kernels of one step = E' x (3 forward + 2 backward SpMM)
            "edge_traversals_per_step": int(e_total * 5),
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
            "loss": round(final_loss, 6), "parity": parity, "configs": other, "minibatch": minib,
            "sage_l3_x64": x64,
        }
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        # leave without tearing NCCL down: destroy_process_group() with captured NCCL kernels still alive
        # in a CUDA graph did not return on the 2-GPU box (round 1); every rank has printed / finished by now
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# --------------------------------------------------------------------------- reference ------
def _cpu_setup(n_timesteps):
    """The reference's CPU path on the first `n_timesteps` timesteps of the same graph."""
    from oracle import pyg_restated as O
    gr = host_graph(1)
    if n_timesteps is not None and n_timesteps < 49:
        keep_n = int((gr.timestep <= n_timesteps).sum())
        ekeep = gr.edge_index[1] < keep_n
        gr.x, gr.y, gr.timestep = gr.x[:keep_n], gr.y[:keep_n], gr.timestep[:keep_n]
        gr.train_mask = gr.train_mask[:keep_n]
        gr.edge_index = gr.edge_index[:, ekeep]
        if gr.train_mask.sum() == 0:
            gr.train_mask = gr.y >= 0
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], dim=1)
    torch.manual_seed(42)
    model = O.build_model(CFG["arch"], gr.x.size(1), CFG)
    opt = torch.optim.Adam(model.parameters(), lr=CFG["lr"], weight_decay=CFG["weight_decay"])
    cw = O.class_weight(gr.y[gr.train_mask])

    def step():
        return O.train_step(model, gr.x, ei, gr.timestep, gr.y, gr.train_mask, cw, opt, CFG["grad_clip"])[0]

    return step, ei.size(1)


def cpu_baseline_sample(budget_s: float):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, edges = _cpu_setup(4)          # calibration on 4 timesteps
    step()
    t0 = time.perf_counter()
    step()
    rate = edges / (time.perf_counter() - t0)
    n_t = max(4, min(49, int(rate * budget_s / 3.0 / (468710 / 49.0))))
    step, edges = _cpu_setup(n_t)
    step()
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    best = min(ts)
    return {"value": round(edges / best / 1e9, 6), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "ms_per_step": round(best * 1e3, 1),
            "sample": f"restated-PyG SAGE-ResBN fp32 train step on the first {n_t}/49 timesteps "
                      f"({edges} edges), 1 warm-up + best of 2",
            "cpu": _cpu_name()}


def _cpu_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, edges = _cpu_setup(4)
    step()
    t0 = time.perf_counter()
    step()
    rate = edges / (time.perf_counter() - t0)
    total = args.steps + args.warmup
    n_t = max(2, min(49, int(rate * 150.0 / total / (468710 / 49.0))))   # whole run <= ~150 s
    step, edges = _cpu_setup(n_t)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ms_step = dt / args.steps * 1e3
    v = round(edges / (ms_step * 1e-3) / 1e9, 6)
    sample = (f"restated-PyG (oracle/pyg_restated.py; torch_geometric is not installable here) SAGE-ResBN fp32 "
              f"train step on the first {n_t}/49 timesteps ({edges} edges) per step")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "rec_k8: SAGE-ResBN 168->64->64->2 (CPU, fp32; reference forces amp off on CPU)",
                       "edges_per_step": edges},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample, "cpu": _cpu_name()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--eager", action="store_true", help="do not capture a CUDA graph (profiling runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--skip-configs", action="store_true", help="skip the other BASELINE configs (profiling runs)")
    ap.add_argument("--skip-x64", action="store_true", help="skip the 64x strong-scaling workload (profiling runs)")
    ap.add_argument("--max-seconds", type=float, default=840.0, help="watchdog: hard-exit if the run hangs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    wd = threading.Timer(args.max_seconds, lambda: (sys.stderr.write("[bench] watchdog: run exceeded "
                         f"{args.max_seconds:.0f} s, exiting\n"), sys.stderr.flush(), os._exit(3)))
    wd.daemon = True
    wd.start()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
