"""BASELINE config 5: 12-layer self-attention encoder / 6-layer decoder, d_model 512, H=8, ~1500-frame utterances.
usage: python tools/run_cfg5.py [B] [band_start] [steps] [profile]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
band = int(sys.argv[2]) if len(sys.argv) > 2 else -100
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
res = bench.cfg5_bench(B=B, band=(band, 0 if band > -1000 else 1600), steps=steps, warmup=3, profile=len(sys.argv) > 4)
print(res)
