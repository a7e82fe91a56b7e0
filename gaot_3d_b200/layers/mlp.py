"""Node-level MLP blocks with the reference's parameter names (state_dict compatible).

`LinearChannelMLP` mirrors reference src/model/layers/mlp.py:308-335 (Linear layers in `fcs`,
exact-erf GELU between layers, none after the last).  `ChannelMLP` mirrors :227-305 (Conv1d k=1 on
channel-first tensors) -- the dataclass default mlp_type='channel'.  Node-level GEMMs stay on
cuBLAS (dense side is a later row); the per-edge kernel MLP never goes through
these modules' forward -- IntegralTransform hands `fcs` weights to the fused CUDA kernel.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class LinearChannelMLP(nn.Module):
    def __init__(self, layers, non_linearity=F.gelu, dropout=0.0):
        super().__init__()
        self.n_layers = len(layers) - 1
        assert self.n_layers >= 1
        self.non_linearity = non_linearity
        self.fcs = nn.ModuleList(nn.Linear(layers[j], layers[j + 1]) for j in range(self.n_layers))
        self.dropout = nn.ModuleList(nn.Dropout(dropout) for _ in range(self.n_layers)) if dropout > 0.0 else None

    def forward(self, x):
        for i, fc in enumerate(self.fcs):
            x = fc(x)
            if i < self.n_layers - 1:
                x = self.non_linearity(x)
            if self.dropout is not None:
                x = self.dropout[i](x)
        return x


class ChannelMLP(nn.Module):
    """Pointwise (kernel-size-1 Conv1d) MLP over [C, N] or [B, C, N...] tensors."""

    def __init__(self, in_channels, out_channels=None, hidden_channels=None, n_layers=2, n_dim=2,
                 non_linearity=F.gelu, dropout=0.0, **kwargs):
        super().__init__()
        self.n_layers = n_layers
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        self.hidden_channels = in_channels if hidden_channels is None else hidden_channels
        self.non_linearity = non_linearity
        self.dropout = nn.ModuleList(nn.Dropout(dropout) for _ in range(n_layers)) if dropout > 0.0 else None
        self.fcs = nn.ModuleList()
        for i in range(n_layers):
            cin = self.in_channels if i == 0 else self.hidden_channels
            cout = self.out_channels if i == n_layers - 1 else self.hidden_channels
            self.fcs.append(nn.Conv1d(cin, cout, 1))

    def forward(self, x):
        reshaped = False
        size = list(x.shape)
        if x.ndim > 3:
            x = x.reshape((*size[:2], -1))
            reshaped = True
        for i, fc in enumerate(self.fcs):
            x = fc(x)
            if i < self.n_layers - 1:
                x = self.non_linearity(x)
            if self.dropout is not None:
                x = self.dropout[i](x)
        if reshaped:
            x = x.reshape((size[0], self.out_channels, *size[2:]))
        return x
