import sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts")
import bench_configs
out = bench_configs.run_all()
for k, v in out.items():
    print(k, round(v["value"]), round(v["e2e"]), v["status"], {a: round(b, 1) for a, b in v["mean_per_solve"].items()})
