"""Development probe: timeline of the chunked host->device upload vs the per-chunk compute."""
import importlib, sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
spa = importlib.import_module("3dspa_code_b200")
ops = spa.ops
dev = torch.device("cuda")
model = spa.TrackAutoEncoder3D()
variables = model.init(0, {"dino_features": 1, "depth_features": 1})
eng = model.bind(variables, "bf16", dev)
host, noise = bench.synth_clip(100, pinned=True)
keys = ["support_tracks", "support_tracks_visible", "dino_features", "depth_features"]
N, T, chunk = 2048, 150, 256
boundary = host["boundary_frame"].to(dev)
cs = torch.cuda.Stream()
cur = torch.cuda.current_stream()
def run(record):
    cs.wait_stream(cur)
    t0 = torch.cuda.Event(enable_timing=True); t0.record(cur)
    evs = []
    pieces = [(n0, min(n0 + chunk, N)) for n0 in range(0, N, chunk)]
    def upload(p):
        n0, n1 = p
        with torch.cuda.stream(cs):
            dv = {k: host[k][0:1, n0:n1].to(dev, non_blocking=True) for k in keys}
            ev = torch.cuda.Event(enable_timing=True); ev.record(cs)
        return dv, ev
    nxt = upload(pieces[0])
    for i, (n0, n1) in enumerate(pieces):
        dv, ev = nxt
        if i + 1 < len(pieces):
            nxt = upload(pieces[i + 1])
        cur.wait_event(ev)
        for t in dv.values():
            t.record_stream(cur)
        a = torch.cuda.Event(enable_timing=True); a.record(cur)
        x = eng.embed_tracks(dv["support_tracks"], dv["dino_features"], dv["depth_features"], readout=True)
        km = ops.build_key_mask(dv["support_tracks_visible"], boundary[0:1], has_readout=True)
        st = eng.transformer("itt", x, n1 - n0, T + 1, km, out_rows="first")
        b = torch.cuda.Event(enable_timing=True); b.record(cur)
        evs.append((ev, a, b))
    torch.cuda.synchronize()
    if record:
        for i, (ev, a, b) in enumerate(evs):
            print(f"chunk {i}: copy done {t0.elapsed_time(ev):7.2f}  compute start {t0.elapsed_time(a):7.2f}  end {t0.elapsed_time(b):7.2f}")
run(False); run(False)
with torch.no_grad():
    run(True)


# real path, with events injected around the pieces of Engine.encode
import types
orig_transformer = eng.transformer
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.current_stream()); marks.append((name, e))
def tr(short, *a, **k):
    mark(short + " start")
    out = orig_transformer(short, *a, **k)
    mark(short + " end")
    return out
eng.transformer = tr
with torch.no_grad():
    for rep in range(3):
        marks.clear()
        torch.cuda.synchronize()
        mark("t0")
        z = eng.encode(host)
        mark("encode end")
        torch.cuda.synchronize()
    t0 = marks[0][1]
    for name, e in marks[1:]:
        print(f"{name:12s} {t0.elapsed_time(e):8.2f}")
