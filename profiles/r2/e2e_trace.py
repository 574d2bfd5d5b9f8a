"""Host-buffer (e2e) steps of config 2 with the API / launch timeline on stderr (PSKMER_TRACE=1).
   PSKMER_TRACE=1 [PSKMER_SC1=lean] python profiles/r2/e2e_trace.py [config]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from phenotypeseeker_b200.pipeline import KmerAssociation
from phenotypeseeker_b200 import synth_gpu
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
plan = synth_gpu.config_plan(cfg)
dev = torch.device("cuda", 0)
r = synth_gpu.Renderer(plan, dev)
text, spans = r.render(range(plan.n_samples))
host = torch.empty(text.numel(), dtype=torch.uint8).pin_memory(); host.copy_(text); hv = host.numpy()
bufs = [hv[o:o + n] for s, (o, n) in sorted(spans.items())]
del text; torch.cuda.empty_cache()
ka = KmerAssociation(device=0)
N = plan.n_samples
kw = dict(min_samples=2, max_samples=N - 2, pvalue_cutoff=0.05, omit_b=False)
for it in range(4):
    if it == 3:
        ka.ctx.profile_reset(); ka.ctx.profile(True)
    torch.cuda.synchronize(); t0 = time.time()
    ka.count(bufs, 16, 1); ka.build(); res = ka.test(plan.pheno, plan.binary, plan.weights, **kw)
    torch.cuda.synchronize(); print("step", it, round((time.time() - t0) * 1e3, 2), "ms", file=sys.stderr)
ka.ctx.profile(False)
tab = ka.ctx.profile_table()
print({k: round(v["ms"], 2) for k, v in sorted(tab.items(), key=lambda kv: -kv[1]["ms"])[:10]}, file=sys.stderr)
