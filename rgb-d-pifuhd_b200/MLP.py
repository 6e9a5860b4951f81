"""Parameter container with the reference MLP's layout (`MLP.py:12-40`).

Same constructor, same ``filters.{i}.weight|bias`` state_dict keys (so the reference's
checkpoints load), but no per-layer PyTorch forward: the stack is evaluated by the
tcgen05 layer kernel in libpifu_b200.so through ``PIFuNetwNML.query`` / ``PIFuMRNet.query``.
"""
import torch.nn as nn


class MLP(nn.Module):
    def __init__(self, filter_channels, merge_layer=0, res_layers=[], norm="group", last_op=None):
        super().__init__()
        self.filter_channels = list(filter_channels)
        self.res_layers = list(res_layers)
        self.merge_layer = merge_layer if merge_layer > 0 else len(filter_channels) // 2   # MLP.py:25
        self.norm = norm
        self.last_op = last_op
        self.filters = nn.ModuleList()
        self.norms = nn.ModuleList()
        last = len(filter_channels) - 2
        for i in range(len(filter_channels) - 1):
            cin = filter_channels[i] + (filter_channels[0] if i in self.res_layers else 0)
            self.filters.append(nn.Conv1d(cin, filter_channels[i + 1], 1))
            if i != last and norm == "group":
                self.norms.append(nn.GroupNorm(32, filter_channels[i + 1]))
            elif i != last and norm == "batch":
                self.norms.append(nn.BatchNorm1d(filter_channels[i + 1]))

    def forward(self, feature):
        raise NotImplementedError(
            "pifu_b200.MLP holds parameters only; occupancy is evaluated by the fused CUDA path "
            "behind PIFuNetwNML.query / PIFuMRNet.query")
