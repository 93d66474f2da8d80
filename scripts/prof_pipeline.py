"""cProfile of the run_save pipeline (bench_extras.pipeline) on the synthetic store: where does the host time go?"""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, bench_extras as X
from himo_b200 import weights
from himo_b200.engine import SeFlowPPEngine
dev = torch.device("cuda", 0)
eng = SeFlowPPEngine(weights.synth_deflowpp_state_dict(0), device=dev, max_points=bench.N_POINTS)
frames = bench.make_frames(0, 2)
X.pipeline(eng, frames, 0, 0)        # warm
pr = cProfile.Profile()
pr.enable()
out = X.pipeline(eng, frames, 0, 0)
pr.disable()
print({k: out[k] for k in ("save_frames_per_s", "save_seconds", "frames")})
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
