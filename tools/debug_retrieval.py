import numpy as np, torch
from helping_hand_for_egocentric_videos_b200 import ops
from oracle.golden_cases import KNOWN_K, KNOWN_REL, KNOWN_SIM, synth_retrieval
print(ops.retrieval_rows(KNOWN_SIM, KNOWN_REL, 1, KNOWN_K)); torch.cuda.synchronize()
print(ops.retrieval_rows(KNOWN_SIM, KNOWN_REL, 0)); torch.cuda.synchronize()
for N, M in [(5, 7), (40, 300), (8, 2000), (4, 9668)]:
    sim, rel = synth_retrieval(N, M, 1)
    print(N, M, ops.retrieval_rows(sim, rel, 0)[:3]); torch.cuda.synchronize()
    print(N, M, ops.retrieval_rows(sim, rel, 1)[:3]); torch.cuda.synchronize()
