"""A/B timing of the batched SV kernels (GPU box): general register kernel vs lean kernel.

The library reads its selection switches once per process (MBQC_SV_KERNEL_REG, MBQC_LEAN_CTA), so
this script re-executes itself once per variant and prints one JSON line per (variant, pattern,
batch, streams)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {"reg": {"MBQC_SV_KERNEL_REG": "1", "MBQC_JIT": "0"}, "lean128": {"MBQC_LEAN_CTA": "128", "MBQC_JIT": "0"},
            "lean64": {"MBQC_LEAN_CTA": "64", "MBQC_JIT": "0"},
            "jit128": {"MBQC_LEAN_CTA": "128", "MBQC_JIT": "force"}, "jit64": {"MBQC_LEAN_CTA": "64", "MBQC_JIT": "force"},
            # alternative builds of the library (build.sh with MBQC_BUILD_OUT=build/_mbqc_<name>.so)
            "notab": {"MBQC_LIB_PATH": os.path.join(ROOT, "build", "_mbqc_notab.so"), "MBQC_LEAN_CTA": "128"},
            "alt": {"MBQC_LIB_PATH": os.path.join(ROOT, "build", "_mbqc_alt.so")}}
CASES = [("grid_cluster", [2, 6]), ("linear_cluster", [5]), ("grid_cluster", [3, 5]), ("grid_cluster", [4, 5])]
if os.environ.get("PERF_CASES"):  # e.g. PERF_CASES=0,1
    CASES = [CASES[int(i)] for i in os.environ["PERF_CASES"].split(",")]
STREAMS = [int(x) for x in os.environ.get("PERF_STREAM_LIST", "1,4").split(",")]


def child(variant):
    from scripts.perf_sv import time_kernel
    for spec in CASES:
        for B in (65536, 1 << 20, 1 << 22):
            if spec[1] == [4, 5] and B > (1 << 20):
                continue
            for S in STREAMS:
                if S > 1 and B > 65536:
                    continue
                os.environ["PERF_STREAMS"] = str(S)
                us = time_kernel(spec, B, reps=400 if B <= 65536 else 40)
                print(json.dumps({"variant": variant, "pattern": f"{spec[0]}{tuple(spec[1])}", "B": B, "streams": S,
                                  "us_per_launch": round(us, 3), "G_evals_per_s": round(B / us / 1e3, 3)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        for name in (sys.argv[1:] or list(VARIANTS)):
            env = dict(os.environ, **VARIANTS[name])
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name], env=env, check=False)
