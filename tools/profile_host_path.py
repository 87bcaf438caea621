#!/usr/bin/env python
"""cProfile of the host side of SAETrainer.train_step at the YAML batch (128 rows): at this size the
GPU floor of the graphed step is below the Python cost of one call.   python tools/profile_host_path.py"""
import cProfile
import pstats
import sys
import tempfile
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whisper_sae_b200.config import TrainingConfig  # noqa: E402
from whisper_sae_b200.sae import SAETrainer, TopKSAE  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(42)
tr = SAETrainer(TopKSAE(384, 3072, k=32), TrainingConfig(batch_size=B, use_amp=True, num_workers=0),
                device="cuda:0", run_dir=Path(tempfile.mkdtemp()))
tr.setup_scheduler(100_000)
xs = [torch.randn(B, 384).cuda() for _ in range(8)]
for i in range(50):
    tr.train_step(xs[i % 8])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(2000):
    tr.train_step(xs[i % 8])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
