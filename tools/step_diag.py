import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from onepose_b200 import GATsSuperGlue, synthetic, _lib
lib = _lib.load()
dev = torch.device('cuda', 0)
hp = dict(synthetic.DEFAULT_HPARAMS); sd = synthetic.make_state_dict(0)
model = GATsSuperGlue(hp).eval(); model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}); model = model.to(dev)
B, N, M = 32, 1024, 7000
db, leaves = synthetic.make_object(0, M, 8)
q = [torch.from_numpy(np.stack([synthetic.make_frame(1000 * s + f, db, N)[0] for f in range(B)], 0)).to(dev) for s in range(4)]
model.set_object(torch.from_numpy(db).to(dev), torch.from_numpy(leaves).to(dev), reserve=(B, N))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def run(tag, fn, steps=12, do_flush=True):
    for w in range(3): fn(w % 4)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.perf_counter()
    for s in range(steps):
        if do_flush: flush.fill_(s & 255)
        ev[s][0].record(); o = fn(s % 4); ev[s][1].record()
    torch.cuda.synchronize()
    print(f"{tag:34s} wall/step {1e3*(time.perf_counter()-t0)/steps:6.2f}  dev " + " ".join(f"{a.elapsed_time(b):5.2f}" for a, b in ev), flush=True)
conf_fixed = torch.empty(B, N, M, device=dev)
m0 = torch.empty(B, N, dtype=torch.int64, device=dev); m1 = torch.empty(B, M, dtype=torch.int64, device=dev)
s0 = torch.empty(B, N, device=dev); s1 = torch.empty(B, M, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
def fixed(i, conf=True):
    _lib.check(lib.opb_forward(model._handle, q[i].data_ptr(), None, B, N, m0.data_ptr(), m1.data_ptr(), s0.data_ptr(), s1.data_ptr(), conf_fixed.data_ptr() if conf else None, st), model._handle)
for rep in range(2):
    run("match_frames fresh conf", lambda i: model.match_frames(q[i]))
    run("match_frames no conf", lambda i: model.match_frames(q[i], return_conf=False))
    run("opb_forward fixed conf buffer", lambda i: fixed(i))
    run("opb_forward fixed, no flush", lambda i: fixed(i), do_flush=False)
for name, setter in (("pdl off", lambda v: lib.opb_debug_set_pdl(v)),):
    setter(0); run(name, lambda i: fixed(i)); setter(1); run(name.replace("off", "on"), lambda i: fixed(i))
lib.opb_debug_set_kv_passes(model._handle, 3); run("kv 3 passes", lambda i: fixed(i)); lib.opb_debug_set_kv_passes(model._handle, 2); run("kv 2 passes", lambda i: fixed(i))
import os
os.environ["OPB_PROFILE_DUMP"] = "1"
model.set_profiling(True); fixed(0); model.get_profile(); model.set_profiling(False)
