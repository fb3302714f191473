"""cProfile of the host side of the end-to-end mesh step of scripts/e2e_host_profile.py (200 steps, sorted by own time): which python
frames the ~0.2 ms in front of mvr_mesh_forward are made of.  usage: python scripts/e2e_host_cprofile.py [list]   (list: the python list of per-object CPU meshes instead of the collated batch)"""
import cProfile, io, os, pstats, sys
here = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(here, "e2e_host_profile.py")).read().split("for _ in range(10):")[0]
g = {"__file__": os.path.join(here, "e2e_host_profile.py"), "__name__": "prof"}
exec(compile(src, "e2e_host_profile.py", "exec"), g)
step = g["step"]
for _ in range(20):
    step(None)
pr = cProfile.Profile(); pr.enable()
for _ in range(200):
    step(None)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(32); print(s.getvalue()[:7000])
for n, v in g["log"].items():
    import statistics
    print("%-28s called at %8.1f us, host %6.1f us" % (n, statistics.median(x[0] for x in v[-200:]), statistics.median(x[1] for x in v[-200:])))
