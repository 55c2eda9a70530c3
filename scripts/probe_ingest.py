"""Ingest throughput (files/s): device-side zstd decode vs libzstd on the host. usage: probe_ingest.py [n_distinct] [tile]"""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from sgic_b200 import faiss_compat as faiss
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 10
print(json.dumps(bench.ingest_extra(faiss, torch.device("cuda", 0), nd, tile), indent=1))
