"""Parameter containers with torch-identical names / shapes / initialisation, whose math runs in
libtacorl_b200.so.  They never call torch.nn.functional compute ops."""
import math

import torch
import torch.nn as nn

from .. import ops


class Linear(nn.Module):
    """nn.Linear parameters (weight (out,in), bias (out)); forward = tacorl_gemm with fused bias/act."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features) if self.in_features > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x, act=None):
        return ops.linear(x, self.weight, self.bias, act)


class Conv2dParams(nn.Module):
    """nn.Conv2d parameters only (weight (out,in,k,k), bias); consumed by the fused encoder kernels."""

    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = kernel_size, stride
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(in_channels * kernel_size * kernel_size)
        nn.init.uniform_(self.bias, -bound, bound)


class Marker(nn.Module):
    """Parameter-free placeholder that keeps nn.Sequential indices equal to the reference's
    (activations, Dropout(0), Flatten).  The op itself is fused into the neighbouring kernel."""

    def __init__(self, what=""):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class ReluRNN(nn.Module):
    """nn.RNN(nonlinearity='relu', batch_first=True) parameters + tacorl_rnn_layer kernels."""

    def __init__(self, input_size, hidden_size, num_layers=1, bidirectional=False, dropout=0.0):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("RNN dropout > 0 is not used by the reference configs")
        self.input_size, self.hidden_size = input_size, hidden_size
        self.num_layers, self.bidirectional = num_layers, bidirectional
        D = 2 if bidirectional else 1
        bound = 1.0 / math.sqrt(hidden_size)
        self._names = []
        for l in range(num_layers):
            for d in range(D):
                suf = "_reverse" if d == 1 else ""
                i = input_size if l == 0 else hidden_size * D
                for name, shape in ((f"weight_ih_l{l}{suf}", (hidden_size, i)),
                                    (f"weight_hh_l{l}{suf}", (hidden_size, hidden_size)),
                                    (f"bias_ih_l{l}{suf}", (hidden_size,)),
                                    (f"bias_hh_l{l}{suf}", (hidden_size,))):
                    p = nn.Parameter(torch.empty(*shape).uniform_(-bound, bound))
                    self.register_parameter(name, p)
                    self._names.append(name)

    def weights(self):
        return [getattr(self, n) for n in self._names]

    def forward(self, x, h0=None, last_only=False):
        """x: (B,T,I) batch-first.  Returns (out (B,T,D*H), h_n) or (out[:, -1], None) if last_only."""
        x_tm = x.transpose(0, 1)
        out, hn = ops.relu_rnn(x_tm, self.weights(), self.num_layers, self.bidirectional, last_only, h0)
        if last_only:
            return out, None
        return out.transpose(0, 1), hn

    def forward_time_major(self, x_tm, h0=None, last_only=False, grad_rows=None):
        """grad_rows: only the first `grad_rows` batch rows will ever receive a gradient (see ops.ReluRNN2Fn)."""
        return ops.relu_rnn(x_tm, self.weights(), self.num_layers, self.bidirectional, last_only, h0, grad_rows)
