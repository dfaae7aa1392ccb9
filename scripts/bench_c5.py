"""The C5 leg of bench.py alone (200 k mixed sizes, fp32): python scripts/bench_c5.py   [DAQP_B200_TEAM=0 for a warp per problem]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--c5-problems", type=int, default=200_000)
args = ap.parse_args()
out = bench.leg_c5(torch.device("cuda:0"), args)
out["team_env"] = os.environ.get("DAQP_B200_TEAM", "default")
print(json.dumps(out))
