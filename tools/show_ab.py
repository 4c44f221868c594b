"""Print the per-stage times of tools/ab_probe.py results side by side:  python tools/show_ab.py gpurun_out/r2_x*.json"""
import json, sys
for fn in sys.argv[1:]:
    try:
        d = json.load(open(fn))
    except Exception as e:
        print(fn, "unreadable", e); continue
    print(f"== {fn} [{d.get('tag')}]")
    for n, w in d["workloads"].items():
        for m in ("exact", "fast"):
            if m not in w: continue
            r = w[m]
            print(f"  {n:18s} {m:5s} clear {r.get('ms_clear', 0)*1e3:6.1f} vert {r['ms_vertex']*1e3:6.1f} setup {r['ms_setup']*1e3:6.1f} raster {r['ms_raster']*1e3:6.1f} "
                  f"frag {r['ms_fragment']*1e3:6.1f} post {r['ms_post']*1e3:6.1f} | total {r['ms_total']*1e3:7.1f} graph {r['graph_ms']*1e3:7.1f} us" + (f"  pipe {w['pipe_fps']:.0f} fps" if m == 'fast' and w.get('pipe_fps') else ""))
