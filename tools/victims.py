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


class DGCNNVictim(nn.Module):
    """DGCNN classifier with the layer shapes of the reference's model/dgcnn.py:262-328 (four edge
    convolutions 6->64, 128->64, 128->128, 256->256 with BN + LeakyReLU(0.2) and max over k, 512->emb
    point-wise layer, max+avg pooling, 2 emb->512->256->classes head), written from the public
    architecture.  `graph_feature(x[B,C,N], k) -> [B,2C,N,k]` is injected: this package's
    dgcnn.get_graph_feature or the reference's torch formulation (torch_graph_feature below)."""

    def __init__(self, graph_feature, k=20, emb_dims=1024, num_classes=106, dropout=0.5):
        super().__init__()
        self.gf, self.k = graph_feature, k
        act = lambda: nn.LeakyReLU(negative_slope=0.2)
        self.conv1 = nn.Sequential(nn.Conv2d(6, 64, 1, bias=False), nn.BatchNorm2d(64), act())
        self.conv2 = nn.Sequential(nn.Conv2d(128, 64, 1, bias=False), nn.BatchNorm2d(64), act())
        self.conv3 = nn.Sequential(nn.Conv2d(128, 128, 1, bias=False), nn.BatchNorm2d(128), act())
        self.conv4 = nn.Sequential(nn.Conv2d(256, 256, 1, bias=False), nn.BatchNorm2d(256), act())
        self.conv5 = nn.Sequential(nn.Conv1d(512, emb_dims, 1, bias=False), nn.BatchNorm1d(emb_dims), act())
        self.linear1 = nn.Linear(emb_dims * 2, 512, bias=False)
        self.bn6, self.dp1 = nn.BatchNorm1d(512), nn.Dropout(dropout)
        self.linear2 = nn.Linear(512, 256)
        self.bn7, self.dp2 = nn.BatchNorm1d(256), nn.Dropout(dropout)
        self.linear3 = nn.Linear(256, num_classes)

    def forward(self, x):
        B = x.size(0)
        x1 = self.conv1(self.gf(x, self.k)).max(dim=-1)[0]
        x2 = self.conv2(self.gf(x1, self.k)).max(dim=-1)[0]
        x3 = self.conv3(self.gf(x2, self.k)).max(dim=-1)[0]
        x4 = self.conv4(self.gf(x3, self.k)).max(dim=-1)[0]
        x = self.conv5(torch.cat((x1, x2, x3, x4), dim=1))
        x = torch.cat((F.adaptive_max_pool1d(x, 1).view(B, -1), F.adaptive_avg_pool1d(x, 1).view(B, -1)), 1)
        x = self.dp1(F.leaky_relu(self.bn6(self.linear1(x)), negative_slope=0.2))
        x = self.dp2(F.leaky_relu(self.bn7(self.linear2(x)), negative_slope=0.2))
        x = F.log_softmax(self.linear3(x), -1)
        return x, x, x


def torch_graph_feature(x, k):
    """The reference's formulation (model/dgcnn.py:194-227) with x.device instead of the hard-coded cuda:0."""
    B, C, N = x.shape
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    idx = (-xx - inner - xx.transpose(2, 1)).topk(k=k, dim=-1)[1]
    idx = (idx + torch.arange(0, B, device=x.device).view(-1, 1, 1) * N).view(-1)
    xt = x.transpose(2, 1).contiguous()
    feature = xt.view(B * N, -1)[idx, :].view(B, N, k, C)
    xr = xt.view(B, N, 1, C).repeat(1, 1, k, 1)
    return torch.cat((feature - xr, xr), dim=3).permute(0, 3, 1, 2).contiguous()


def torch_sample_and_group(npoint, radius, nsample, xyz, points):
    """The reference's formulation of model/pointnet2_utils.py:59-135 (FPS Python loop, square_distance + full sort
    ball query, advanced-indexing gathers) on xyz.device -- the baseline the FPS / ball-query kernels replace."""
    B, N, C = xyz.shape
    dev = xyz.device
    cent = torch.zeros(B, npoint, dtype=torch.long, device=dev)
    dist = torch.ones(B, N, device=dev) * 1e10
    far = torch.randint(0, N, (B,), dtype=torch.long).to(dev)
    bi = torch.arange(B, dtype=torch.long, device=dev)
    for i in range(npoint):
        cent[:, i] = far
        c = xyz[bi, far, :].view(B, 1, 3)
        d = torch.sum((xyz - c) ** 2, -1)
        m = d < dist
        dist[m] = d[m]
        far = torch.max(dist, -1)[1]
    new_xyz = xyz[bi[:, None], cent]
    sq = -2 * torch.matmul(new_xyz, xyz.permute(0, 2, 1))
    sq += torch.sum(new_xyz ** 2, -1).view(B, npoint, 1)
    sq += torch.sum(xyz ** 2, -1).view(B, 1, N)
    gi = torch.arange(N, dtype=torch.long, device=dev).view(1, 1, N).repeat([B, npoint, 1])
    gi[sq > radius ** 2] = N
    gi = gi.sort(dim=-1)[0][:, :, :nsample]
    first = gi[:, :, 0].view(B, npoint, 1).repeat([1, 1, nsample])
    mask = gi == N
    gi[mask] = first[mask]
    grouped = xyz[bi[:, None, None], gi] - new_xyz.view(B, npoint, 1, C)
    if points is not None:
        grouped = torch.cat([grouped, points[bi[:, None, None], gi]], dim=-1)
    return new_xyz, grouped


class _SetAbstraction(nn.Module):
    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all, sample_and_group):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all, self.sag = npoint, radius, nsample, group_all, sample_and_group
        self.convs, self.bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel
        for out in mlp:
            self.convs.append(nn.Conv2d(last, out, 1)); self.bns.append(nn.BatchNorm2d(out)); last = out

    def forward(self, xyz, points):
        xyz = xyz.permute(0, 2, 1)
        if points is not None:
            points = points.permute(0, 2, 1)
        if self.group_all:
            B, N, C = xyz.shape
            new_xyz = torch.zeros(B, 1, C, device=xyz.device)
            new_points = xyz.view(B, 1, N, C) if points is None else torch.cat([xyz.view(B, 1, N, C), points.view(B, 1, N, -1)], dim=-1)
        else:
            new_xyz, new_points = self.sag(self.npoint, self.radius, self.nsample, xyz, points)
        new_points = new_points.permute(0, 3, 2, 1)
        for conv, bn in zip(self.convs, self.bns):
            new_points = F.relu(bn(conv(new_points)))
        return new_xyz.permute(0, 2, 1), torch.max(new_points, 2)[0]


class PointNet2SSGVictim(nn.Module):
    """PointNet++ SSG classifier with the layer shapes of the reference's model/pointnet2_SSG.py:230-254
    (SA 512/0.2/32 [64,64,128], SA 128/0.4/64 [128,128,256], global SA [256,512,1024], 1024-512-256-classes),
    written from the public architecture; `sample_and_group(npoint, radius, nsample, xyz, points)` is injected."""

    def __init__(self, sample_and_group, num_classes=106):
        super().__init__()
        self.sa1 = _SetAbstraction(512, 0.2, 32, 3, [64, 64, 128], False, sample_and_group)
        self.sa2 = _SetAbstraction(128, 0.4, 64, 128 + 3, [128, 128, 256], False, sample_and_group)
        self.sa3 = _SetAbstraction(None, None, None, 256 + 3, [256, 512, 1024], True, sample_and_group)
        self.fc1, self.bn1, self.drop1 = nn.Linear(1024, 512), nn.BatchNorm1d(512), nn.Dropout(0.4)
        self.fc2, self.bn2, self.drop2 = nn.Linear(512, 256), nn.BatchNorm1d(256), nn.Dropout(0.4)
        self.fc3 = nn.Linear(256, num_classes)

    def forward(self, xyz):
        B = xyz.shape[0]
        l1_xyz, l1_points = self.sa1(xyz, None)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        _, l3_points = self.sa3(l2_xyz, l2_points)
        x = l3_points.view(B, 1024)
        x = self.drop1(F.relu(self.bn1(self.fc1(x))))
        x = self.drop2(F.relu(self.bn2(self.fc2(x))))
        x = F.log_softmax(self.fc3(x), -1)
        return x, x, x
