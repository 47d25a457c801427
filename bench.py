#!/usr/bin/env python
"""Contract benchmark: GAT layer fwd+bwd edges/sec at the ogbn-proteins shape.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one forward + backward of the GATConv sparse section (logits -> leaky_relu
-> edge-softmax -> dropout -> SpMM -> scaling) over the whole
synthetic graph; dense fc projections are excluded from `value` (they stay cuBLAS) and
included in `e2e`, which goes through the reference-facing `GATConv.forward(graph,
feat_src, feat_edge)` with HOST input buffers.

N > 1: the same graph is 1-D partitioned on destination rows (strong scaling); every
rank all-gathers the source table halo, runs its rows, and the gradient halo is
reduce-scattered back (bot_b200/partition.py).

--impl reference: the oracle port of the reference math (pure torch, CPU) on a bounded
sample of the same workload, on the host cores of this box (DGL itself is not
installable offline: DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# The contract is ONE JSON line on stdout, but libraries print there too (NCCL's version banner: this image sets
# NCCL_DEBUG=VERSION).  main() points file descriptor 1 at stderr for the whole run; the JSON line goes to the real stdout.
_REAL_STDOUT = None


def claim_stdout():
    """Called by main() only (importing this module for its constants must not touch the caller's stdout)."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


import torch  # noqa: E402

# ogbn-proteins layer shape
N_NODES, N_EDGES, HEADS, HID, EDGE_EMB = 132534, 39561252, 6, 80, 16
EDGE_DROP, SLOPE = 0.1, 0.2
METRIC = "GAT layer fwd+bwd edges/sec (ogbn-proteins shape)"
CPU_SAMPLE_EDGES = N_EDGES // 4   # bounded CPU sample: ~4.6 s per fwd+bwd on the GPU box's 16 host cores


def workload_name(n_nodes=N_NODES, n_edges=N_EDGES):
    return (f"ogbn-proteins-shape GATConv layer: N={n_nodes} E={n_edges} H={HEADS} D={HID} fp32, attn_dst + "
            f"edge-feature logits ({EDGE_EMB}-dim edge emb), edge_drop={EDGE_DROP} (training), uniform random edges")


def algorithmic_bytes(E, N_s, N_d, H, D, er=True, ee=True):
    """SURVEY.md section 8(d) gather/streaming model.  Returns (fwd, bwd) bytes per layer pass."""
    R, h = 4 * H * D, 4 * H
    x = (h + 4) if ee else 0
    e_r = 1 if er else 0
    fwd = E * (4 + h + R + x) + N_d * (8 + h * e_r + R + 2 * h)
    bwd = E * ((4 + h + R + x + h) + (8 + 2 * h + R)) + N_d * (8 + 2 * R + 2 * h + 2 * h * e_r) + N_s * (8 + R + h)
    return fwd, bwd


# ----------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self, timeout=3.0):
        # nvidia-smi needs a few hundred ms before its first line: do not start the timed region without it
        t = time.time()
        while self.proc is not None and not self.lines and time.time() - t < timeout:
            time.sleep(0.01)
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # the sampler runs from before the warm-up; only samples that arrived inside the timed region count
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1]
        window = "timed region"
        if not inside:
            inside, window = [ln for (_, ln) in self.lines], "whole run (no sample fell inside the timed region)"
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ----------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------
def synth_edges(n_nodes, n_edges, device, seed=0, power_law=0.0):
    """SURVEY.md section 8d generator: uniform random COO; power_law > 0 draws dst proportional to rank^-power_law."""
    g = torch.Generator(device=device).manual_seed(seed)
    src = torch.randint(0, n_nodes, (n_edges,), device=device, generator=g)
    if power_law > 0:
        w = torch.arange(1, n_nodes + 1, device=device, dtype=torch.float64) ** (-power_law)
        cdf = torch.cumsum(w / w.sum(), 0)
        u = torch.rand(n_edges, device=device, dtype=torch.float64, generator=g)
        dst = torch.searchsorted(cdf, u).clamp(max=n_nodes - 1)
    else:
        dst = torch.randint(0, n_nodes, (n_edges,), device=device, generator=g)
    return src, dst


# ----------------------------------------------------------------------------
# CPU arm: the oracle port of the reference math on a bounded sample
# ----------------------------------------------------------------------------
def cpu_port_run(steps, warmup, n_edges=CPU_SAMPLE_EDGES):
    """fwd+bwd of the same layer math on the host (oracle port, non-materialising form).
    Returns dict(value, unit, cores, kind, sample, ms_per_step)."""
    from oracle import gat_ref

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    src, dst = synth_edges(N_NODES, n_edges, "cpu", seed=0)
    bg = gat_ref.BigGraph(src, dst, N_NODES, N_NODES)
    g = torch.Generator().manual_seed(1)
    ft = torch.randn(N_NODES, HEADS, HID, generator=g)
    el = torch.randn(N_NODES, HEADS, generator=g)
    er = torch.randn(N_NODES, HEADS, generator=g)
    ee = torch.randn(n_edges, HEADS, generator=g)
    gout = torch.randn(N_NODES, HEADS, HID, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out, a, z = gat_ref.gat_sparse_big_forward(bg, ft, el, er, ee, SLOPE)
            gat_ref.gat_sparse_big_backward(bg, ft, a, z, out, gout, SLOPE)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return {"value": n_edges / t, "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"same layer math on a {n_edges}-edge uniform subsample ({N_EDGES // n_edges}x fewer edges) over all {N_NODES} "
                      f"nodes, fwd+bwd, {len(times)} timed steps, oracle/gat_ref.py gat_sparse_big_* "
                      f"(segment_reduce + per-head sparse CSR matmul), torch {torch.__version__} CPU",
            "ms_per_step": t * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_port_run(args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "edges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "sampled_edges_per_step": CPU_SAMPLE_EDGES},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    import bot_b200
    from bot_b200 import _lib, functional
    from bot_b200.functional import gat_fused
    from bot_b200.ogbn_proteins import GATConv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device: bot_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_nodes, n_edges = args.nodes, args.edges
    src, dst = synth_edges(n_nodes, n_edges, dev, seed=0)
    gen = torch.Generator(device=dev).manual_seed(1)

    if world == 1:
        graph = bot_b200.Graph(src, dst, n_nodes)
        graph.create_formats_()
        layer = None
    else:
        from bot_b200 import partition

        layer = partition.PartitionedGraph(src, dst, n_nodes)
        graph = layer.local
        graph.create_formats_()
    del src, dst
    E_local = graph.number_of_edges()
    n_src_l, n_dst_l = graph.number_of_src_nodes(), graph.number_of_dst_nodes()

    # ---------------- device-resident leg (value) ----------------
    n_own = n_dst_l if world > 1 else n_nodes
    ft_own = torch.randn(n_own, HEADS, HID, device=dev, generator=gen).requires_grad_(True)
    el_own = torch.randn(n_own, HEADS, device=dev, generator=gen).requires_grad_(True)
    er = torch.randn(n_dst_l, HEADS, device=dev, generator=gen).requires_grad_(True)
    # edge logits as the drop-in GATConv emits them: (E, pad_heads(H)) = (E, 8), one 32-byte record per edge
    ee = torch.randn(E_local, functional.pad_heads(HEADS), device=dev, generator=gen).requires_grad_(True)
    gout = torch.randn(n_dst_l, HEADS, HID, device=dev, generator=gen)

    seeds = iter(range(1, 1 << 30))

    def step_resident():
        # the reference's draw: exactly int(E * p) uniformly random edges dropped (models.py:136-139)
        keep = functional.edge_drop_keep(E_local, int(E_local * EDGE_DROP), next(seeds), dev)
        if world == 1:
            out = gat_fused(graph, ft_own, el_own, er, ee, keep, None, None, None, SLOPE, 0.0, 0)
        else:
            out = layer.gat(ft_own, el_own, er, ee, keep, None, None, None, SLOPE, 0.0, 0)
        out.backward(gout)
        for t in (ft_own, el_own, er, ee):
            t.grad = None

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    launches0 = lib.botgat_launch_count()
    kt = functional.KernelTimer()
    functional.timer = kt
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    sampler.mark_end()
    functional.timer = None
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.botgat_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_edges / (ms_step * 1e-3)
    ktot = kt.totals()

    # roofline of the dominant kernel (largest share of the timed region)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"
    bf, bb = algorithmic_bytes(E_local, n_src_l, n_dst_l, HEADS, HID)
    # backward bytes split by pass (first bracket = dst pass over the in-CSR, second = src pass)
    R, h = 4 * HEADS * HID, 4 * HEADS
    b_dst = E_local * (4 + h + R + (h + 4) + h) + n_dst_l * (8 + R + 2 * h + 2 * h)
    b_src = bb - b_dst
    # the dst-major re-gather of the reference's backward (b_dst) is eliminated algorithmically: the src pass
    # produces the per-edge dot products from the rows it gathers anyway.  Per-kernel rooflines count only
    # the bytes of the kernel's own gather; `whole_step` keeps SURVEY's full fwd+bwd figure.
    alg = {"gat_fwd": bf, "gat_bwd_src": b_src}
    kernels = {}
    for name, (n, tot) in ktot.items():
        avg = tot / n
        k = {"launches": n, "avg_ms": round(avg, 4), "share_of_step": round(tot / ms_total, 4)}
        if name in alg:
            k["algorithmic_bytes"] = alg[name]
            k["achieved_GBs"] = round(alg[name] / (avg * 1e-3) / 1e9, 1)
            k["frac_of_peak"] = round(k["achieved_GBs"] / peak, 4)
        kernels[name] = k
    dom = max((n for n in kernels if n in alg), key=lambda n: kernels[n]["avg_ms"] * kernels[n]["launches"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_GBs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_of_peak"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom],
                "whole_step": {"algorithmic_bytes": bf + bb,
                               "achieved_GBs": round((bf + bb) / (ms_step * 1e-3) / 1e9, 1),
                               "frac": round((bf + bb) / (ms_step * 1e-3) / 1e9 / peak, 4)}}

    # second roofline: the feature slabs are L2-resident by construction, so the gathers are
    # bounded by L2 -> SM bandwidth, measured on this pool with tools/l2peak.cu (profiles/l2_peak.json)
    l2path = os.path.join(ROOT, "profiles", "l2_peak.json")
    if os.path.exists(l2path):
        l2peak = float(json.load(open(l2path))["l2_read_gbs"])
        roofline["l2"] = {"bound": "l2", "peak": l2peak, "unit": "GB/s", "peak_source": "profiles/l2_peak.json (measured)",
                          "kernels": {n: {"achieved": kernels[n]["achieved_GBs"],
                                          "frac": round(kernels[n]["achieved_GBs"] / l2peak, 4)} for n in alg if n in kernels}}
    roofline["note"] = ("achieved = SURVEY 8d gather-model bytes / CUDA-event time; it exceeds the HBM peak because each "
                        "(N x D) head slab stays L2-resident: `traffic` is the DRAM bytes ncu measured for the same launch")

    del ft_own, el_own, er, ee, gout
    torch.cuda.empty_cache()

    # ---------------- end-to-end leg through GATConv.forward with host buffers ----------------
    e2e = None
    if world == 1:
        torch.manual_seed(0)
        conv = GATConv(HEADS * HID, EDGE_EMB, HID, n_heads=HEADS, edge_drop=EDGE_DROP).to(dev)
        conv.train()
        h_host = torch.randn(n_nodes, HEADS * HID).pin_memory()
        fe_host = torch.randn(n_edges, EDGE_EMB).pin_memory()
        h2d = h_host.numel() * 4 + fe_host.numel() * 4

        copy_stream = torch.cuda.Stream()

        def finish(y):
            loss = y.square().mean()
            loss.backward()
            conv.zero_grad(set_to_none=True)
            return float(loss.item())  # device -> host read of the step's result

        def step_serial():
            # copy and layer of the same step back to back (how the reference's loop is written): node features first
            # (the projections need them at once), edge features on a copy stream so that their 2.5 GB transfer
            # overlaps the node-side GEMMs and the edge-drop draw (bot_b200.Deferred)
            h = h_host.to(dev, non_blocking=True).requires_grad_(True)
            with torch.cuda.stream(copy_stream):
                fe = fe_host.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return finish(conv(graph, h, bot_b200.Deferred(fe, ev, requires_grad=True)))

        def run_serial(k):
            for _ in range(k):
                step_serial()

        def run_fed(k):
            # bot_b200.HostFeed: the copies of step i+1 are enqueued before step i's layer and run beside it
            # (two device buffer sets); every step still copies its own 2.8 GB and reads its loss back
            feed = bot_b200.HostFeed(dev, depth=2)
            feed.submit(h_host, fe_host)
            for i in range(k):
                if i + 1 < k:
                    feed.submit(h_host, fe_host)
                h, fe = feed.take(requires_grad=(0, 1))
                finish(conv(graph, h, fe))

        def timed(run, k):
            torch.cuda.synchronize()
            e0.record()
            run(k)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k

        w_e2e = max(1, min(args.warmup, 3))
        run_serial(w_e2e)
        k_serial = max(1, min(args.steps, 5))
        ms_serial = timed(run_serial, k_serial)
        torch.cuda.empty_cache()
        run_fed(w_e2e)
        k_e2e = max(1, args.steps)
        ms_e2e = timed(run_fed, k_e2e)
        e2e = {"value": n_edges / (ms_e2e * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "ms_per_step": round(ms_e2e, 3), "steps": k_e2e,
               "call": "bot_b200.ogbn_proteins.GATConv.forward(graph, feat_src, feat_edge) + backward; feat_src (N,480) and "
                       "feat_edge (E,16) copied from pinned host memory every step through bot_b200.HostFeed (double-buffered: "
                       "step i+1's copy is enqueued before step i's layer; the first copy and the last layer of the timed "
                       "region are not overlapped), loss read back every step; graph structure resident (the reference moves "
                       "the graph once, run.py:539); includes the five nn.Linear projections",
               "unpipelined": {"ms_per_step": round(ms_serial, 3), "value": n_edges / (ms_serial * 1e-3), "steps": k_serial,
                               "note": "same call with each step's copy issued inside the step (no prefetch)"}}
        del conv, h_host, fe_host
        torch.cuda.empty_cache()
    else:
        # the same layer call as at N=1, partitioned: every rank feeds ITS shard — the node features of its owned rows and
        # the edge features of its local edges — from pinned host memory through bot_b200.HostFeed, runs the node-side
        # projections on its rows, the partitioned sparse section (halo all-gather / reduce-scatter inside), all-reduces
        # the weight gradients and reads its loss back
        from bot_b200.functional import edge_logits

        torch.manual_seed(0)
        conv = GATConv(HEADS * HID, EDGE_EMB, HID, n_heads=HEADS, edge_drop=EDGE_DROP).to(dev)   # same weights everywhere
        conv.train()
        host_shard = [torch.randn(n_own, HEADS * HID).pin_memory(), torch.randn(E_local, EDGE_EMB).pin_memory()]
        h2d_rank = sum(t.numel() * 4 for t in host_shard)
        params = [p for p in conv.parameters()]

        def run_fed_ranks(k):
            feed = bot_b200.HostFeed(dev, depth=2)
            feed.submit(*host_shard)
            for i in range(k):
                if i + 1 < k:
                    feed.submit(*host_shard)
                x, fe = (d.wait() for d in feed.take(requires_grad=(0, 1)))
                ft = conv.src_fc(x).view(-1, HEADS, HID)
                resid = conv.dst_fc(x).view(-1, HEADS, HID)
                el, er = conv.attn_src_fc(x), conv.attn_dst_fc(x)
                ee = edge_logits(fe, conv.attn_edge_fc.weight)
                keep = functional.edge_drop_keep(E_local, int(E_local * EDGE_DROP), next(seeds), dev)
                y = layer.gat(ft, el, er, ee, keep, None, None, None, SLOPE, 0.0, 0) + resid
                loss = y.square().mean()
                loss.backward()
                flat = torch.cat([p.grad.flatten() for p in params])
                dist.all_reduce(flat)                      # data-parallel weight gradients
                conv.zero_grad(set_to_none=True)
                float(loss.item())  # device -> host read of the step's result

        run_fed_ranks(max(1, min(args.warmup, 3)))
        k_e2e = max(1, args.steps)
        barrier()
        e0.record()
        run_fed_ranks(k_e2e)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1), float(h2d_rank)], device=dev, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_e2e = float(tmax[0].item()) / k_e2e
        e2e = {"value": n_edges / (ms_e2e * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": int(t[1].item()),
               "d2h_bytes_per_step": 4 * world, "ms_per_step": round(ms_e2e, 3), "steps": k_e2e,
               "call": "the N=1 layer, partitioned: on every rank nn.Linear projections of its owned rows + "
                       "bot_b200.partition.PartitionedGraph.gat + residual + backward + all-reduce of the weight gradients; each "
                       "rank copies its shard (feat_src of its rows, feat_edge of its local edges) from pinned host memory every "
                       "step through bot_b200.HostFeed (step i+1's copy enqueued before step i's layer) and reads its loss back; "
                       "max over ranks; h2d bytes summed over ranks"}
        del host_shard, conv

    # secondary number: the same layer on a heavy-tailed graph (dst ~ rank^-0.8; the hottest row has ~2 % of all edges)
    skew = None
    if world == 1 and not args.no_skew:
        s2, d2 = synth_edges(n_nodes, n_edges, dev, seed=0, power_law=0.8)
        g2 = bot_b200.Graph(s2, d2, n_nodes)
        g2.create_formats_()
        del s2, d2
        gen2 = torch.Generator(device=dev).manual_seed(2)
        t_ft = torch.randn(n_nodes, HEADS, HID, device=dev, generator=gen2).requires_grad_(True)
        t_el = torch.randn(n_nodes, HEADS, device=dev, generator=gen2).requires_grad_(True)
        t_er = torch.randn(n_nodes, HEADS, device=dev, generator=gen2).requires_grad_(True)
        t_ee = torch.randn(n_edges, functional.pad_heads(HEADS), device=dev, generator=gen2).requires_grad_(True)
        t_go = torch.randn(n_nodes, HEADS, HID, device=dev, generator=gen2)

        def step_skew():
            keep = functional.edge_drop_keep(n_edges, int(n_edges * EDGE_DROP), next(seeds), dev)
            gat_fused(g2, t_ft, t_el, t_er, t_ee, keep, None, None, None, SLOPE, 0.0, 0).backward(t_go)
            for t in (t_ft, t_el, t_er, t_ee):
                t.grad = None

        for _ in range(2):
            step_skew()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            step_skew()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        skew = {"dst_distribution": "rank^-0.8", "max_in_degree": int(g2._info.max_in_deg), "ms_per_step": round(ms, 3),
                "value": n_edges / (ms * 1e-3), "unit": "edges/s", "split_row_slots": int(g2._info.n_slots_in)}
        del g2, t_ft, t_el, t_er, t_ee, t_go
        torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_port_run(steps=2, warmup=1)
        cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n_nodes, n_edges), "l2": "inputs_exceed_l2 (no flush between iterations: "
                       "the per-step working set, >= 2.5 GB, is 20x the 126 MB L2)",
                       "parallelism": "single GPU" if world == 1 else f"1-D dst-row partition x{world}, halo all-gather + "
                       "gradient reduce-scatter (NCCL)",
                       "timed": "edge-drop mask draw (exactly int(E*p) edges, botgat_edge_drop_draw) + edge staging + fused fwd + bwd (node/src/dst passes) + edge unstage"},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "e2e": e2e, "skew_variant": skew,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=N_NODES)
    ap.add_argument("--edges", type=int, default=N_EDGES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-skew", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
