"""A tiny point-cloud classifier used as the victim of the loop-level parity tests (shared by
oracle/make_golden.py, which runs the UNMODIFIED reference attack loops against it on CPU, and by
the GPU tests, which run this repository's device-resident loops against the same weights).
forward(x[B,3,K]) -> (logits[B,classes], None, None), the calling convention of every reference victim."""
import numpy as np
import torch
import torch.nn as nn


class TinyVictim(nn.Module):
    def __init__(self, classes=7):
        super().__init__()
        self.c1 = nn.Conv1d(3, 16, 1)
        self.c2 = nn.Conv1d(16, 32, 1)
        self.fc = nn.Linear(32, classes)

    def forward(self, x):
        h = torch.relu(self.c1(x))
        h = torch.relu(self.c2(h))
        return self.fc(h.max(dim=2)[0]), None, None


def make(seed=0, classes=7):
    torch.manual_seed(seed)
    m = TinyVictim(classes)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(3.0)                      # sharper logits: the attack has something to do
    return m.eval()


def state_to_npz(model):
    return {"victim__" + k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def from_npz(g, classes=7):
    m = TinyVictim(classes)
    m.load_state_dict({k[len("victim__"):]: torch.from_numpy(np.asarray(g[k])) for k in (g.files if hasattr(g, "files") else g.keys()) if k.startswith("victim__")})
    return m.eval()
