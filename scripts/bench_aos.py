"""The array-of-struct leg of bench.py alone (daqp_quadprog_batch on C3 problems in pageable host memory):
python scripts/bench_aos.py [--aos-problems N]   (DAQP_B200_AOS_PIECE_MB tunes the piece size)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from daqp_b200.problems import generate_g1

ap = argparse.ArgumentParser(); ap.add_argument("--aos-problems", type=int, default=20000)
args = ap.parse_args()
b = generate_g1(args.aos_problems, 50, 150, 0, 40, seed=3)
hn = {"H": b.H, "f": b.f, "A": b.A, "bupper": b.bupper, "blower": b.blower}
aos, lat = bench.leg_aos_latency(hn, 50, 150, 0, args)
print(json.dumps({"piece_mb": os.environ.get("DAQP_B200_AOS_PIECE_MB", "default"), "aos": aos["value"], "problems": aos["problems"], "latency_us": lat["median_us"]}))
