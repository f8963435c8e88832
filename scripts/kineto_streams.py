"""CUPTI timeline of the live training step, per stream and per phase (geometry / forward / backward / optimizer): which
stream bounds the backward, how long each phase spans, where the GPU idles.  python scripts/kineto_streams.py [steps]"""
import collections, json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from sparse_rcnn_b200 import pipeline, scn
dev = torch.device("cuda:0"); scn.set_precision("tf32")
tr = pipeline.BackboneTrainer(dev)
batches = [bench.make_inputs(s) for s in range(4)]
batches = [((d[0].pin_memory(), d[1].pin_memory(), d[2], d[3], d[4]), l.pin_memory()) for d, l in batches]
tr.stage_uploads = True
tr.build_late = os.environ.get("BUILD_LATE", "1") == "1"
for i in range(8): tr.step(*batches[i % 4], next_batch=batches[(i + 1) % 4])
torch.cuda.synchronize()
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(NS): tr.step(*batches[i % 4], next_batch=batches[(i + 1) % 4])
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "scn_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
def union(iv):
    iv = sorted(iv); tot = 0.0; cs, ce = iv[0]
    for s, e in iv[1:]:
        if s > ce: tot += ce - cs; cs, ce = s, e
        else: ce = max(ce, e)
    return tot + ce - cs
GEO = ("k_hash", "k_scan", "k_radix", "k_tile_book", "k_subm_map", "k_stride", "k_strided", "k_input_rule", "k_morton", "k_level", "k_compact", "k_seg")
def phase_of(name, seen_ce_bwd, seen_adam):
    return None
# split into steps at the optimizer kernel
steps, cur = [], []
for e in ev:
    cur.append(e)
    if "FusedOptimizer" in e["name"] or "multi_tensor_apply" in e["name"]:
        if len(cur) < 10 and steps: steps[-1].extend(cur)      # second flat buffer of the same optimizer step
        else: steps.append(cur)
        cur = []
print("steps found: %d (kernels per step: %s)" % (len(steps), [len(s) for s in steps]))
for si, st in enumerate(steps[1:], 1):      # the first one starts mid-way (pack_all of the previous step)
    t0 = st[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in st)
    by_stream = collections.defaultdict(list)
    for e in st: by_stream[e["args"].get("stream")].append((e["ts"], e["ts"] + e["dur"]))
    # phases: up to the first conv kernel = geometry; up to k_ce_bwd = forward; rest = backward (+ optimizer)
    geo = [e for e in st if any(g in e["name"] for g in GEO)]
    print("   geometry kernels: %d, busy %.3f ms, from +%.3f to +%.3f ms of the step" % (len(geo), union([(e["ts"], e["ts"] + e["dur"]) for e in geo]) / 1e3,
          (geo[0]["ts"] - st[0]["ts"]) / 1e3, (geo[-1]["ts"] + geo[-1]["dur"] - st[0]["ts"]) / 1e3))
    i_conv = next(i for i, e in enumerate(st) if "k_conv" in e["name"])
    i_ce = next(i for i, e in enumerate(st) if "k_ce_bwd" in e["name"])
    tg, tf = st[i_conv]["ts"], st[i_ce]["ts"]
    print("step %d: span %.3f ms | head (pack, H2D, rulebooks) %.3f | forward %.3f | backward+opt %.3f | GPU busy (any stream) %.3f" % (
        si, (t1 - t0) / 1e3, (tg - t0) / 1e3, (tf - tg) / 1e3, (t1 - tf) / 1e3, union([(e["ts"], e["ts"] + e["dur"]) for e in st]) / 1e3))
    for s, iv in sorted(by_stream.items(), key=lambda x: -len(x[1])):
        bw = [(a, b) for a, b in iv if a >= tf]
        print("   stream %s: %d launches, busy %.3f ms (in backward: %d launches, busy %.3f ms, last end +%.3f ms)" % (
            s, len(iv), union(iv) / 1e3, len(bw), union(bw) / 1e3 if bw else 0.0, ((max(b for _, b in bw) - tf) / 1e3) if bw else 0.0))
    if si == 1:
        for nm, lo, hi in (("head", t0, tg), ("forward", tg, tf), ("backward", tf, t1)):
            seg = [e for e in st if lo <= e["ts"] < hi]
            busy = union([(e["ts"], e["ts"] + e["dur"]) for e in seg]) if seg else 0
            agg = collections.defaultdict(lambda: [0, 0.0])
            for e in seg: agg[e["name"].split("(")[0][:48]][0] += 1; agg[e["name"].split("(")[0][:48]][1] += e["dur"]
            print("   -- %s: %d launches, busy %.3f of %.3f ms; top: %s" % (nm, len(seg), busy / 1e3, (hi - lo) / 1e3, "; ".join(
                "%s x%d %.0fus" % (n, c, t) for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:8])))
            # idle gaps inside the phase
            end = seg[0]["ts"]; gaps = []
            for e in seg:
                if e["ts"] - end > 4: gaps.append((e["ts"] - end, e["name"].split("(")[0][:32]))
                end = max(end, e["ts"] + e["dur"])
            print("      idle gaps > 4 us: %d, total %.3f ms; largest: %s" % (len(gaps), sum(g for g, _ in gaps) / 1e3,
                  ", ".join("%.0f us before %s" % g for g in sorted(gaps, reverse=True)[:8])))

# effective cost of every kernel of the first steady step's forward chain: start-to-start intervals on the main stream
st = steps[1]
i_conv = next(i for i, e in enumerate(st) if "k_conv" in e["name"])
i_ce = next(i for i, e in enumerate(st) if "k_ce_bwd" in e["name"])
main = max(collections.Counter(e["args"].get("stream") for e in st).items(), key=lambda x: x[1])[0]
chain = [e for e in st[i_conv:i_ce] if e["args"].get("stream") == main]
print("forward chain on stream %s: kernel, grid, duration us, start-to-next-start us" % main)
for a, b in zip(chain, chain[1:] + [None]):
    nxt = (b["ts"] - a["ts"]) if b is not None else a["dur"]
    print("   %-36s grid %5s  dur %6.1f  step %6.1f" % (a["name"].split("(")[0].replace("void scn::", "").replace("scn::", "")[:34],
                                                       a["args"].get("grid", ["?"])[0], a["dur"], nxt))
