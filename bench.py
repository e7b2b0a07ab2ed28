#!/usr/bin/env python
"""bench.py — graphs/sec of the SignNet hot path (forward + backward) on ZINC-shaped synthetic batches.

    python bench.py --gpus N --steps K --warmup W            # B200 path (this repo), one rank per GPU under torchrun
    python bench.py --impl reference --steps K --warmup W    # the reference's arithmetic on the host CPU (oracle port)

One JSON line on stdout (rank 0).  A "step" = zero_grad -> SignNetGNN(data) -> L1 loss -> backward on one batch of
`--batch` graphs per GPU (weak scaling); with N > 1 the flat fp32 gradient buffer is all-reduced over NCCL every step
(bucketed, overlapped with the backward) and the line also carries a `strong` sub-record: BASELINE.json configs[3] as
written, ONE global batch of `--batch` graphs sharded over the N GPUs (ddp.shard_batch).
  value        device-resident: inputs already in HBM, CUDA events around K steps, max over ranks
  e2e          through the public module API from pinned HOST buffers: H2D of the batch + step + D2H of the loss
  roofline     phi GIN-aggregate kernel (K1): algorithmic bytes / CUDA-event duration vs measured HBM peak
  cpu_baseline oracle port (oracle/restate.py) timed on the host cores on the SAME batch (one step; bounded sample)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "graphs/sec (SignNet fwd+bwd, ZINC-shape batch)"
WORKLOAD = ("cfg4: ZINC-shape graphs (n~23.2, E~2.1N), SignNetGNN(PyG ZINC tree) n_hid=128 k=N_max(37) masked, "
            "nl_signnet=8, nl_rho=1 (SetTransformer), nl_gnn=6 GINE, n_out=1; optimizer step excluded")
CFG = dict(n_hid=128, n_out=1, nl_signnet=8, nl_gnn=6, flavour="zinc")


def make_batch(B, seed):
    from signnet_basisnet_b200.synth import synth_batch

    d = synth_batch(B, "zinc", seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    d.y = torch.randn(B, CFG["n_out"], generator=g)
    return d


def workload_config(batch, world, nodes, edges):
    """The `config` object both arms print (the reference arm runs rank 0's batch of the N = 1 job)."""
    return {"workload": WORKLOAD, "graphs_per_gpu": batch, "global_batch": batch * world, "nodes_per_gpu": nodes,
            "edges_per_gpu": edges, "parallelism": f"dp{world}",
            "l2": "working set per step (>= 0.6 GB per activation tensor) exceeds the 126 MB L2",
            "weights": "random init (reference default init, seed 0)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None     # the timed region (rows outside it are dropped): nvidia-smi is started BEFORE the
        # warm-up, because its start-up (process + NVML initialisation) perturbs a launch-bound step for ~100 ms

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def __enter__(self):
        if os.environ.get("SB_NO_CLOCKS") == "1":
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            now = time.time()
            if self.t0 is not None and now >= self.t0 and (self.t1 is None or now <= self.t1 + 0.11):
                self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist

    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200.ddp import FlatGradAllReduce
    from signnet_basisnet_b200.sign_net import SignNetGNN

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; see --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    torch.manual_seed(0)
    model = SignNetGNN(None, None, CFG["n_hid"], CFG["n_out"], CFG["nl_signnet"], CFG["nl_gnn"],
                       flavour=CFG["flavour"]).to(dev).train()
    sync = FlatGradAllReduce(model, world) if world > 1 else None
    if sync is not None:
        sync.broadcast_parameters()

    host = make_batch(args.batch, seed=1000 + rank).pin_memory()
    resident = host.to(dev)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.__dict__.values() if torch.is_tensor(v))
    params = list(model.parameters())
    use_sync = [True]

    from signnet_basisnet_b200.layout import pad4, prepare_batch

    # the loader's stream runs at high priority: its few small kernels (H2D copy, integer bookkeeping, one 32-byte read-back
    # the host waits for) otherwise queue behind a launch queue that is several persistent 148-CTA kernels deep
    copy_stream = torch.cuda.Stream(device=dev, priority=-1)
    LD = pad4(CFG["n_hid"])

    def twin(d):
        """Two views of one resident batch: every step is treated as a NEW batch (its bookkeeping - graph offsets, CSR,
        slot layout, one device->host size read - is rebuilt inside the timed region), prepared on the side stream one
        step ahead of the compute stream, exactly as the e2e loop does for the batch it has just copied."""
        a, b = type(d)(**d.__dict__), type(d)(**d.__dict__)
        a.__dict__.pop("_b200_graph_index", None), b.__dict__.pop("_b200_graph_index", None)
        return [a, b]

    def step(data, nxt=None):
        for p in params:
            p.grad = None
        if nxt is not None:
            prepare_batch(nxt, LD, stream=copy_stream)       # bookkeeping of the NEXT step, off the critical path
        if getattr(data, "_b200_graph_index", None) is None:
            prepare_batch(data, LD)                          # first step only
        out = model(data)
        data.__dict__.pop("_b200_graph_index", None)

        loss = (out - data.y).abs().mean()
        loss.backward()
        if sync is not None and use_sync[0]:
            sync.allreduce()
        return loss

    # End to end: every step copies ITS batch host -> device from pinned memory (on a copy stream, issued one step ahead
    # like a prefetching DataLoader would, so the transfer overlaps the previous step's backward) and reads the loss back.
    LOSS_LAG, LOSS_RING = 3, 4
    loss_pin = torch.empty(LOSS_RING, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(blocking=True) for _ in range(LOSS_RING)]
    e2e_state = {"i": 0, "losses": [], "read": 0, "split": []}

    def fetch():
        with torch.cuda.stream(copy_stream):
            d = host.to(dev, non_blocking=True)
        prepare_batch(d, LD, stream=copy_stream)   # the loader's side: H2D copy + bookkeeping of the batch it delivers
        return d

    # The loader runs in its own thread, like a DataLoader worker: the H2D copy, the integer bookkeeping and the one
    # 32-byte read-back the host has to wait for (graph / slot counts) happen two batches ahead of the step that uses them,
    # so the launching thread never sits in that wait (on some boxes of the pool it took 30-60 ms per batch for several
    # steps in a row and the launch queue ran dry; SB_BENCH_DEBUG=1 prints the per-section host times).
    import queue
    import threading

    ready = queue.Queue(maxsize=2)
    loader_stop = threading.Event()

    def loader_main():
        torch.cuda.set_device(local)
        while not loader_stop.is_set():
            d = fetch()
            while not loader_stop.is_set():
                try:
                    ready.put(d, timeout=0.05)
                    break
                except queue.Full:
                    pass

    loader = threading.Thread(target=loader_main, daemon=True)

    def step_e2e():
        t0_ = time.perf_counter()
        if not loader.is_alive():
            loader.start()
        data = ready.get()
        for v in data.__dict__.values():
            if torch.is_tensor(v):
                v.record_stream(torch.cuda.current_stream())
        t1_ = time.perf_counter()
        # D2H read of the loss, every step, without draining the launch pipeline: the scalar goes to pinned host memory
        # with an asynchronous copy and is consumed LOSS_LAG steps later, once its event has completed (what a training
        # loop that logs the loss does when `.item()` must not stall the next step's launches; a host-side hiccup of the
        # shared box then no longer idles the GPU).  Every loss of the timed region is read before its closing barrier.
        loss = step(data)
        i = e2e_state["i"]
        e2e_state["i"] = i + 1
        loss_pin[i % LOSS_RING].copy_(loss.detach(), non_blocking=True)
        loss_ev[i % LOSS_RING].record()
        t2_ = time.perf_counter()
        if i >= LOSS_LAG:
            drain_loss(i - LOSS_LAG)
        e2e_state["split"].append((t1_ - t0_, t2_ - t1_, time.perf_counter() - t2_))

    def drain_loss(j):
        loss_ev[j % LOSS_RING].synchronize()
        e2e_state["losses"].append(float(loss_pin[j % LOSS_RING]))
        e2e_state["read"] = j + 1

    def drain_all():
        for j in range(e2e_state.get("read", 0), e2e_state["i"]):
            drain_loss(j)

    ring = twin(resident)
    pos = [0]

    def step_resident():
        a, b = ring[pos[0] & 1], ring[(pos[0] + 1) & 1]
        pos[0] += 1
        return step(a, b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import gc

    def timed(fn, n, after=None):
        # the cyclic collector is paused inside a timed region (a full collection of this process is a 10-20 ms pause
        # that lands in one random step of a 20-step window); reference cycles wait until the region ends
        gc.collect()
        gc_was = gc.isenabled()
        if os.environ.get("SB_BENCH_GC") != "1":
            gc.disable()
        try:
            return _timed(fn, n, after)
        finally:
            if gc_was:
                gc.enable()

    def _timed(fn, n, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dbg = os.environ.get("SB_BENCH_DEBUG") == "1"
        tt = []
        for _ in range(n):
            t_ = time.perf_counter()
            fn()
            tt.append(time.perf_counter() - t_)
        if after is not None:
            after()   # still inside the timed region (e.g. the e2e loop's outstanding loss reads)
        e1.record()
        if dbg:
            st_ = torch.cuda.memory_stats()
            sys.stderr.write("cpu ms per step: " + " ".join(f"{1e3 * t:.1f}" for t in tt) + "\n")
            sys.stderr.write("allocator: " + " ".join(f"{k}={st_.get(k)}" for k in (
                "num_alloc_retries", "segment.all.allocated", "segment.all.freed", "reserved_bytes.all.current",
                "reserved_bytes.all.peak", "allocated_bytes.all.peak")) + "\n")
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    with ClockSampler(local) as clk:
        for _ in range(max(args.warmup, 3)):
            step_resident()
        c0 = _lib.launch_count
        clk.mark_start()
        ms = timed(step_resident, args.steps)
        clk.mark_stop()
    launches = _lib.launch_count - c0
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps, after=drain_all)
    loader_stop.set()
    loader.join(timeout=5)
    if os.environ.get("SB_BENCH_DEBUG") == "1":
        sys.stderr.write("e2e host ms per step (fetch | step | loss wait): " + " ".join(
            f"{1e3 * a_:.0f}|{1e3 * b_:.0f}|{1e3 * c_:.0f}" for a_, b_, c_ in e2e_state["split"][-args.steps:]) + "\n")

    # N > 1: (a) what the gradient exchange costs after overlap = step time with it minus step time without it;
    # (b) strong scaling, BASELINE.json configs[3] as written: ONE global batch of --batch graphs sharded over the GPUs
    comm = strong = None
    if world > 1:
        use_sync[0] = False
        ms_local = timed(step_resident, args.steps)
        use_sync[0] = True
        comm = {"collective": "ncclAllReduce(avg) of one flat fp32 gradient buffer in buckets on a side stream",
                "payload_bytes": sync.last_payload_bytes, "bucket_bytes": sync.bucket_bytes(),
                "exposed_ms_per_step": round((ms - ms_local) / args.steps, 3),
                "ms_per_step_without_exchange": round(ms_local / args.steps, 3)}
        from signnet_basisnet_b200.ddp import shard_batch

        shard = shard_batch(make_batch(args.batch, seed=1000), world, rank).to(dev)
        ring[:] = twin(shard)
        for _ in range(max(args.warmup, 3)):
            step_resident()
        ms_strong = timed(step_resident, args.steps)
        ring[:] = twin(resident)
        strong = {"scaling": "strong", "global_batch": args.batch, "graphs_per_gpu": shard.num_graphs,
                  "value": round(args.batch * args.steps / (ms_strong * 1e-3), 1), "unit": "graphs/s",
                  "ms_per_step": round(ms_strong / args.steps, 3),
                  "note": "same model, same step; rank r runs graphs [r*B/N, (r+1)*B/N) of ONE seeded batch"}

    # per-entry-point breakdown with CUDA events on the launching stream (separate pass, not part of `value`)
    _lib.profile_start()
    for _ in range(args.steps):
        step_resident()
    prof = _lib.profile_stop()
    total_prof = sum(t for _, t in prof.values()) or 1.0
    kernels = [{"entry": k, "launches_per_step": c / args.steps, "ms_per_step": t / args.steps,
                "share": t / total_prof} for k, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])][:14]

    # roofline of the phi aggregate (K1).  SURVEY §8d: bytes = 2*4*ld*S*R + 16*E per launch (read X, write A, read the
    # int64 edge_index once).  In the default path the forward aggregate runs INSIDE gin_lin_fused_kernel, which also
    # writes the first Linear's output H (one more activation tensor): its algorithmic bytes are 3*4*ld*S*R + 16*E.
    # The stand-alone K1 kernel still runs in the backward (transposed CSR; dA, G, X in, G out = 4 tensors) and is also
    # timed here on its own, live, with CUDA events on the launching stream.
    gi = prepare_batch(resident, LD)
    sl = gi.slots_all(CFG["n_hid"])
    ld, T1 = CFG["n_hid"], 4 * CFG["n_hid"] * 2 * sl.R
    peak, peak_src = peaks()
    traffic = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "agg_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic_src = "static: " + tj.get("source", "ncu --set full capture of an earlier run (profiles/)")

    def rec(tag, nbytes):
        if tag not in prof:
            return None
        c, t = prof[tag]
        ach = nbytes / (t / c * 1e-3) / 1e9
        return {"achieved": round(ach, 1), "frac": round(ach / peak, 4), "bytes_per_launch": nbytes,
                "avg_launch_us": round(t / c * 1e3, 1), "launches_per_step": c / args.steps}

    from signnet_basisnet_b200.phi import gin_agg

    xs = torch.randn(2, sl.R, ld, device=dev)
    ys = torch.empty_like(xs)
    for _ in range(3):
        gin_agg(xs, ys, sl, 2, ld)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(7):
        gin_agg(xs, ys, sl, 2, ld)
    ev1.record()
    torch.cuda.synchronize()
    t_alone = ev0.elapsed_time(ev1) / 7
    alone_bytes = 2 * T1 + 16 * gi.E
    alone = {"kernel": "gin_agg_tma_kernel (K1 alone, forward CSR), 7 launches timed live outside the step",
             "achieved": round(alone_bytes / (t_alone * 1e-3) / 1e9, 1), "bytes_per_launch": alone_bytes,
             "frac": round(alone_bytes / (t_alone * 1e-3) / 1e9 / peak, 4), "avg_launch_us": round(t_alone * 1e3, 1),
             "traffic": (tj.get("dram_bytes_per_launch") if traffic_src else None), "traffic_source": traffic_src}
    del xs, ys
    # every phi-size streaming / contraction kernel of the step against the same measured HBM peak: algorithmic bytes
    # (activation tensors T1 = 4*ld*S*R bytes each, as DESIGN.md §4 counts them) / CUDA-event time per launch
    rows_phi, Tr = 2 * sl.R, 4 * ld * sl.R
    alg = {f"sb_linear_fwd[K={ld},N={ld},rows={rows_phi}]": (2 * T1, "X in, Y out"),
           f"sb_linear_wgrad[N={ld},K={ld},rows={rows_phi}]": (2 * T1, "dY, X in"),
           "sb_bn_apply_bwd": (3 * T1, "gout, y in, dz out (BatchNorm backward: coefficients + apply)"),
           "sb_bn_bwd_reduce": (2 * T1, "gout, y in (BatchNorm backward, sums)"),
           "sb_bn_apply_fwd": (3 * T1, "y, residual in, x out (BatchNorm forward: coefficients + ReLU + residual; 2 tensors in layer 0)"),
           "sb_gin_linear_fused_fwd": (3 * T1 + 16 * gi.E, "X in, A and H out"),
           f"sb_gin_agg[ld={ld},bwd]": (4 * T1 + 16 * gi.E, "dA, G, X in, G out"),
           "sb_attention_fwd": (4 * Tr, "q, k, v in, o out"),
           "sb_attention_bwd": (7 * Tr, "q, k, v, do in, dq, dk, dv out"),
           f"sb_linear_fwd[K={ld},N={ld},rows={sl.R}]": (2 * Tr, "X in, Y out (rho)"),
           f"sb_linear_wgrad[N={ld},K={ld},rows={sl.R}]": (2 * Tr, "dY, X in (rho)")}
    for kr in kernels:
        if kr["entry"] in alg and kr["launches_per_step"] > 0:
            nb, what = alg[kr["entry"]]
            us = kr["ms_per_step"] / kr["launches_per_step"] * 1e3
            kr.update({"avg_launch_us": round(us, 1), "algorithmic_bytes_per_launch": nb, "algorithmic_bytes": what,
                       "hbm_frac": round(nb / (us * 1e-6) / 1e9 / peak, 3)})
    fused = rec("sb_gin_linear_fused_fwd", 3 * T1 + 16 * gi.E)
    fwd = rec(f"sb_gin_agg[ld={ld}]", alone_bytes)
    bwd = rec(f"sb_gin_agg[ld={ld},bwd]", 4 * T1 + 16 * gi.E)
    main = fused or fwd
    roof = None
    if main is not None:
        roof = {"kernel": ("gin_lin_fused_kernel (phi aggregate K1 + first Linear of the MaskedMLP, forward; csrc/gin_lin_fused.cu)"
                           if fused else "gin_agg_tma_kernel (phi aggregate, forward)"),
                "bound": "hbm", "achieved": main["achieved"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": main["frac"], "traffic": None if fused else alone["traffic"],
                "traffic_source": None if fused else traffic_src,
                "algorithmic_bytes_per_launch": main["bytes_per_launch"],
                "algorithmic_bytes_note": ("X read + A written + H written (3 activation tensors) + int64 edge_index once; "
                                           "K1 alone is 2 tensors (SURVEY §8d)" if fused else "SURVEY §8d"),
                "avg_launch_us": main["avg_launch_us"], "launches_per_step": main["launches_per_step"],
                "slot_rows_R": sl.R, "dense_slot_bytes_per_launch": 2 * 4 * ld * 2 * gi.N * sl.k + 16 * gi.E,
                "aggregate_alone": alone, "backward": bwd}
        fp = os.path.join(ROOT, "profiles", "fused_traffic.json")
        if fused and os.path.exists(fp):
            fj = json.load(open(fp))
            roof["traffic"], roof["traffic_source"] = fj.get("dram_bytes_per_launch"), "static: " + fj.get("source", "")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    graphs = args.batch * world * args.steps
    line = {
        "metric": METRIC, "value": round(graphs / (ms * 1e-3), 1), "unit": "graphs/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args.batch, world, gi.N, gi.E),
        "e2e": {"value": round(graphs / (ms_e2e * 1e-3), 1), "unit": "graphs/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3),
                "loss_read": "every step, async copy to pinned memory, consumed 3 steps later (all inside the timed region)",
                "losses_read": len(e2e_state["losses"])},
        "gpu_launches": launches, "gpu_launches_note": "C-ABI entry-point calls inside the timed region (each "
                                                        "enqueues >= 1 kernel of libsignnet_b200.so)",
        "clocks": clk.summary(), "roofline": roof, "kernels": kernels,
    }
    if comm is not None:
        line["grad_exchange"] = comm
        line["strong"] = strong
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.batch, args.cpu_sample)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- CPU arm (oracle port)
def _oracle_step_fn(B, seed=1000):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import restate  # the CPU restatement of the reference's arithmetic (test infrastructure, used here as the baseline)

    from signnet_basisnet_b200.sign_net import SignNetGNN

    torch.manual_seed(0)
    model = SignNetGNN(None, None, CFG["n_hid"], CFG["n_out"], CFG["nl_signnet"], CFG["nl_gnn"], flavour=CFG["flavour"])
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    d = make_batch(B, seed)

    def step():
        for v in sd.values():
            if v.requires_grad:
                v.grad = None
        out = restate.sign_net_gnn(d, sd, CFG["nl_signnet"], CFG["nl_gnn"], nl_rho=1, ignore_eigval=True, training=True,
                                   attn_dropout=0.1)
        loss = (out - d.y).abs().mean()
        loss.backward()
        return float(loss.detach())

    return step


def _time_steps(step, n):
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    return time.perf_counter() - t0


def _phi_only_step_fn(B, seed=1000):
    """phi(+v) + phi(-v) forward + backward alone (BASELINE.md §3 asks for whole-step AND phi-only graphs/s)."""
    import restate

    from signnet_basisnet_b200.sign_net import SignNetGNN

    torch.manual_seed(0)
    model = SignNetGNN(None, None, CFG["n_hid"], CFG["n_out"], CFG["nl_signnet"], CFG["nl_gnn"], flavour=CFG["flavour"])
    sd = {k[len("sign_net.phi."):]: v.detach().clone() for k, v in model.state_dict().items()
          if k.startswith("sign_net.phi.")}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    d = make_batch(B, seed)
    _, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
    mask = restate.slot_mask(d.batch, eigV.shape[1])

    def step():
        for v in sd.values():
            if v.requires_grad:
                v.grad = None
        restate.phi_pm(eigV, d.edge_index, mask, sd, "", CFG["nl_signnet"], True).sum().backward()

    return step


def cpu_baseline(B, B_small):
    """Oracle port on the host cores, same seeded batch as the GPU arm (rank 0): warm-up on a small batch, then ONE
    timed fwd+bwd step on the B graphs (about 15 s on 16 cores - the bounded sample of the contract)."""
    torch.set_num_threads(os.cpu_count() or 1)
    _oracle_step_fn(B_small)()          # warm-up: thread pool, allocator, autograd graph code paths
    dt = _time_steps(_oracle_step_fn(B), 1)
    return {"value": round(B / dt, 2), "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
            "graphs_per_step": B,
            "sample": f"1 warm-up step on {B_small} graphs + 1 timed fwd+bwd step of the same model on the same {B} "
                      f"ZINC-shape graphs as the GPU arm (oracle/restate.py, torch CPU fp32, {dt:.1f} s of CPU work)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    B = args.batch                      # the SAME 1024-graph batch (seed 1000 = rank 0 of the GPU arm)
    step = _oracle_step_fn(B)
    warm = max(1, min(args.warmup, 1))
    steps = max(1, min(args.steps, 2))  # ~15 s per step on 16 cores: the whole run stays within a few minutes
    _time_steps(step, warm)
    dt = _time_steps(step, steps)
    v = round(B * steps / dt, 2)
    d = make_batch(B, 1000)
    # secondary figures: phi alone on the same batch, and the small-batch throughput round 1 reported
    phi_dt = _time_steps(_phi_only_step_fn(B), 1)
    small = _oracle_step_fn(args.cpu_sample)
    small()
    small_dt = _time_steps(small, 3)
    sample = (f"{steps} timed steps (after {warm} warm-up) of fwd+bwd on the same {B} ZINC-shape graphs as the GPU arm with "
              f"the oracle port of the reference (torch CPU fp32, all host threads)")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "graphs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps, "warmup": warm, "ms_per_step": round(dt / steps * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(B, 1, int(d.batch.numel()), int(d.edge_index.shape[1])),
        "graphs_per_step": B,
        "cpu_baseline": {"value": v, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "graphs_per_step": B, "sample": sample},
        "phi_only": {"value": round(B / phi_dt, 2), "unit": "graphs/s", "ms_per_step": round(phi_dt * 1e3, 1),
                     "what": "phi(+v)+phi(-v) forward+backward alone, same batch (BASELINE.md §3)"},
        "small_batch": {"graphs_per_step": args.cpu_sample, "value": round(args.cpu_sample * 3 / small_dt, 2),
                        "unit": "graphs/s", "what": "the 64-graph figure round 1 reported, kept for continuity"},
        "e2e": {"value": v, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout: library chatter written to fd 1 (e.g. NCCL's version banner) goes to
    stderr instead; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="graphs per GPU")
    ap.add_argument("--cpu-sample", type=int, default=64, help="graphs per step of the CPU warm-up / small-batch figure")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
