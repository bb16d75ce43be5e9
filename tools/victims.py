"""Randomly initialised victim classifiers for benchmarking (no trained weights ship with the
reference).  PointNet classifier with the layer shapes the reference's
model/pointnet.py:130-148 uses (input transform net, 3->64->128->1024 shared MLP, max pool,
1024->512->256->k head), written from the public architecture; forward(x[B,3,N]) returns
(log_probs, trans, None) like every victim of the reference."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _TNet(nn.Module):
    def __init__(self, k=3):
        super().__init__()
        self.k = k
        self.c1, self.c2, self.c3 = nn.Conv1d(k, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, 1024, 1)
        self.f1, self.f2, self.f3 = nn.Linear(1024, 512), nn.Linear(512, 256), nn.Linear(256, k * k)
        self.b1, self.b2, self.b3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(1024)
        self.b4, self.b5 = nn.BatchNorm1d(512), nn.BatchNorm1d(256)

    def forward(self, x):
        B = x.shape[0]
        x = F.relu(self.b1(self.c1(x))); x = F.relu(self.b2(self.c2(x))); x = F.relu(self.b3(self.c3(x)))
        x = torch.max(x, 2)[0]
        x = F.relu(self.b4(self.f1(x))); x = F.relu(self.b5(self.f2(x)))
        x = self.f3(x) + torch.eye(self.k, device=x.device).flatten()
        return x.view(B, self.k, self.k)


class PointNetVictim(nn.Module):
    def __init__(self, num_classes=106):
        super().__init__()
        self.stn = _TNet(3)
        self.c1, self.c2, self.c3 = nn.Conv1d(3, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, 1024, 1)
        self.b1, self.b2, self.b3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(1024)
        self.f1, self.f2, self.f3 = nn.Linear(1024, 512), nn.Linear(512, 256), nn.Linear(256, num_classes)
        self.drop = nn.Dropout(0.3)
        self.b4, self.b5 = nn.BatchNorm1d(512), nn.BatchNorm1d(256)

    def forward(self, x):
        trans = self.stn(x)
        x = torch.bmm(x.transpose(2, 1), trans).transpose(2, 1)
        x = F.relu(self.b1(self.c1(x))); x = F.relu(self.b2(self.c2(x))); x = self.b3(self.c3(x))
        x = torch.max(x, 2)[0]
        x = F.relu(self.b4(self.f1(x))); x = F.relu(self.b5(self.drop(self.f2(x))))
        return F.log_softmax(self.f3(x), dim=1), trans, None
