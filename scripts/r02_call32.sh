set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fastnsf.py tests/test_gpu_cli.py -m gpu -q -x 2>&1 | tail -4
timeout 400 python - <<'PY' 2>&1 | tail -5
import json, torch, bench, bench_extras as X
dev = torch.device("cuda", 0)
fr = bench.make_frames(0, 1)[0]
out = X.fastnsf(dev, fr, 1635.0)
print(json.dumps({k: out[k] for k in ("ms_per_iter", "dt_build_ms", "configured_run", "engine", "engine_stream")}))
import himo_b200.engine as E
for w in (3, 4):
    import time
    eng = E.FastNSFEngine(device=dev, itr_num=5000, early_patience=10, n_workers=w)
    for _ in eng.infer_stream(fr for _ in range(w)): pass
    eng._frame_no = 0
    t0 = time.perf_counter()
    for _ in eng.infer_stream(fr for _ in range(8)): pass
    torch.cuda.synchronize()
    print("workers", w, "ms per pair", (time.perf_counter() - t0) / 8 * 1e3, eng.last_iterations)
PY
