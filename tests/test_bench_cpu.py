"""CPU: the parts of bench.py that do not touch the GPU - how the engine-level trace legs are
launched and read back (one subprocess per arm, tools/engine_trace.py), with the tool replaced by
a stub that records its command line."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STUB = '''
import json, sys
a = sys.argv[1:]
out = a[a.index("--json") + 1]
summary = {"tokens_per_s": 10.0, "p50_latency_s": 1.0, "p90_latency_s": 2.0, "mean_latency_s": 1.5,
           "requests_per_s": 1.0, "num_requests": 7, "finished": 7, "generated_tokens": 70,
           "elapsed_s": 7.0, "steps": 3, "schedule_s": 0.1, "execute_model_s": 6.0,
           "predictor_s": 0.05, "other_host_s": 0.9, "add_requests_s": 0.2, "scores": "synthetic"}
if "--request-rate" in a:
    summary["arrivals"] = "7 requests over 0.2 s"
json.dump({"summary": summary, "argv": a,
           "engine": {"init_s": 1.0, "num_gpu_blocks": 5, "installed": {"scheduler": "LTR tick",
                                                                          "prefill": "skipped: x"}}},
          open(out, "w"))
'''


class _Clocks:
    def summary(self, t0, t1):
        return {"sm_mhz": 0, "window_s": t1 - t0}


def test_trace_legs_command_lines(tmp_path, monkeypatch):
    import bench
    tools = tmp_path / "tools"
    tools.mkdir()
    (tools / "engine_trace.py").write_text(STUB)
    monkeypatch.setattr(bench, "TOOLS", str(tools))
    args = types.SimpleNamespace(trace_model="llama-2-7b", trace_prompts=10000, trace_stop=1000,
                                 trace_poisson_rate=32.0, trace_poisson_time=10.0,
                                 trace_poisson_max_output=192)
    burst = bench.run_trace_arm("b200", args, _Clocks(), 60.0)
    assert burst["tokens_per_s"] == 10.0 and burst["engine_init_s"] == 1.0
    assert burst["installed"] == ["scheduler"] and "arrivals" not in burst
    poisson = bench.run_trace_arm("ref_cuda", args, _Clocks(), 60.0, poisson=True)
    assert poisson["arrivals"] == "7 requests over 0.2 s" and poisson["num_requests"] == 7

    # the command lines: what tools/engine_trace.py is asked to do in either leg
    seen = {}
    real_run = bench.subprocess.run

    def spy(cmd, **kw):
        seen["cmd"] = cmd
        return real_run(cmd, **kw)
    monkeypatch.setattr(bench.subprocess, "run", spy)
    bench.run_trace_arm("b200", args, _Clocks(), 60.0)
    c = seen["cmd"]
    assert c[c.index("--num-prompts") + 1] == "10000" and c[c.index("--stop-after") + 1] == "1000"
    assert "--request-rate" not in c
    bench.run_trace_arm("b200", args, _Clocks(), 60.0, poisson=True)
    c = seen["cmd"]
    assert c[c.index("--stop-after") + 1] == "0" and c[c.index("--request-rate") + 1] == "32.0"
    assert c[c.index("--request-time") + 1] == "10.0" and c[c.index("--max-output") + 1] == "192"
    # enough prompts are generated to fill the arrival window
    assert int(c[c.index("--num-prompts") + 1]) >= 32 * 10


def test_trace_leg_failure_is_reported_not_raised(tmp_path, monkeypatch):
    import bench
    tools = tmp_path / "tools"
    tools.mkdir()
    (tools / "engine_trace.py").write_text("import sys; print('boom'); sys.exit(3)")
    monkeypatch.setattr(bench, "TOOLS", str(tools))
    args = types.SimpleNamespace(trace_model="m", trace_prompts=1, trace_stop=0,
                                 trace_poisson_rate=1.0, trace_poisson_time=1.0,
                                 trace_poisson_max_output=8)
    r = bench.run_trace_arm("b200", args, _Clocks(), 60.0, poisson=True)
    assert r["error"] == "rc=3" and "boom" in r["tail"]
    assert json.dumps(r)
