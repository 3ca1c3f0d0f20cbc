"""Where does the part of a bench step go that is not kernels?  (GPU box; bench.py's Runner at the headline configuration.)

Wall-clock marks on the host around the points where the device is known to be idle: after the one synchronising read of the
rollout (AsyncPPO._worker_fused) until the first optimiser launch is queued, and from the end of learn() (its status read
synchronises) until the next rollout is queued.  Prints the averages over the timed steps.

  python tools/host_gap.py [steps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]] + sys.argv[1:]

import bench  # noqa: E402


def main():
    import torch as t

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    sys.argv = [sys.argv[0]]
    args = bench.parse()
    dev = t.device("cuda", 0)
    t.cuda.set_device(dev)
    flush = t.empty(256 << 20, dtype=t.uint8, device=dev)
    run = bench.Runner(args.cfg, None, dev, flush)
    ap, ppo = run.ap, run.ppo
    from prl_b200 import ops

    marks = {}
    pc = time.perf_counter

    orig_rollout, orig_replay, orig_step = ops.rollout, t.cuda.CUDAGraph.replay, ops.ppo_step_tc
    orig_cpu = t.Tensor.cpu

    evs = []

    def rollout(*a, **k):
        marks.setdefault("rollout_queued", pc())
        e = [t.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        r = orig_rollout(*a, **k)
        e[1].record()
        evs.append(e)
        return r

    def replay(self):
        marks.setdefault("first_update_queue_start", pc())
        r = orig_replay(self)
        marks.setdefault("first_update_queued", pc())
        return r

    ops.rollout = rollout
    t.cuda.CUDAGraph.replay = replay
    import AsyncTools.AsyncPPO as A

    orig_fused = A.AsyncPPO._worker_fused

    def fused(self, initial_states=None):
        r = orig_fused(self, initial_states)
        marks["worker_end"] = pc()
        return r

    A.AsyncPPO._worker_fused = fused
    # the synchronising read inside the worker: time of its return = device idle from here on
    scores_cpu = {"n": 0}

    def cpu(self, *a, **k):
        r = orig_cpu(self, *a, **k)
        if self.dtype == t.float64 and self.numel() == 2 and "rollout_queued" in marks and "sync_returned" not in marks:
            marks["sync_returned"] = pc()
        return r

    t.Tensor.cpu = cpu

    for _ in range(5):
        run.step(False)
    t.cuda.synchronize()
    acc = {}
    t_prev_end = None
    wall0 = pc()
    for _ in range(steps):
        marks.clear()
        s0 = pc()
        run.step(False)   # ends with the status read of learn() (synchronising) + the L2 flush fill (queued)
        s1 = pc()
        d = {
            "step start -> rollout queued (reset, host set-up; device idle after the previous step's flush)": marks["rollout_queued"] - s0,
            "rollout queued -> scores read back (the rollout itself, 6.1 ms, + GAE)": marks["sync_returned"] - marks["rollout_queued"],
            "scores read -> worker() returns (transfer queued)": marks["worker_end"] - marks["sync_returned"],
            "worker() returns -> first graph replay call (learn() prologue; device idle apart from the small kernels)": marks["first_update_queue_start"] - marks["worker_end"],
            "first graph replay call -> returned": marks["first_update_queued"] - marks["first_update_queue_start"],
            "first replay returned -> step() returns (the optimiser launches run)": s1 - marks["first_update_queued"],
            "whole step (wall)": s1 - s0,
        }
        for k, v in d.items():
            acc[k] = acc.get(k, 0.0) + v
    t.cuda.synchronize()
    wall1 = pc()
    print(f"{steps} steps, {1e3 * (wall1 - wall0) / steps:.3f} ms per step (wall)")
    print(f"  {sum(a.elapsed_time(b) for a, b in evs[-steps:]) / steps:8.3f} ms  CUDA events around ops.rollout (the kernel on the device)")
    for k, v in acc.items():
        print(f"  {1e3 * v / steps:8.3f} ms  {k}")
    # learn(), line by line (wall clock between consecutive line events of its own frame; one step)
    from PPO.PPO import PPO as PP

    code = getattr(PP.learn, "__wrapped__", PP.learn).__code__
    lines = []

    def tracer(frame, event, arg):
        if frame.f_code is code:
            def local(frame, event, arg):
                if event == "line":
                    lines.append((frame.f_lineno, pc()))
                return local
            lines.append((frame.f_lineno, pc()))
            return local
        return None

    run.step(False)
    ap.step_score = 0
    ap.reward_score = 0
    ap.worker()
    t_w = pc()
    sys.settrace(tracer)
    ppo.learn()
    sys.settrace(None)
    flush.fill_(1)
    print(f"learn(): first line event {1e3 * (lines[0][1] - t_w):.3f} ms after worker() returned; lines that took > 20 us:")
    for (ln, t0), (_, t1) in zip(lines, lines[1:]):
        if t1 - t0 > 20e-6:
            print(f"    line {ln}: {1e3 * (t1 - t0):.3f} ms")
    # the Python side of one step, function by function (the waits show up under the synchronising reads)
    import cProfile
    import pstats

    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        run.step(False)
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()
