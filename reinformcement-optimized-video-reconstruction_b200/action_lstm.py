"""Drop-in for the reference `action_lstm` module: `ActionLSTM(hidden_dim, num_layers, batch_size)`.

Same constructor, `forward(action, new_tensor)`, `reset_hidden_states()`, parameter names
(`lstm.weight_ih`, `lstm.weight_hh`, `lstm.bias_ih`, `lstm.bias_hh`, `fc.weight`, `fc.bias`) and the
same plain-attribute hidden state (`hx`, `cx` persist across calls and are NOT in state_dict) as
rovr/action_lstm.py:6-42. At the reference's batch size (1) both layers are weight-read-bound
GEMVs (33.3 M fp32 parameters = 133 MB per call), so they run on the fp32 weight-streaming
kernels, with a fused LSTM-cell pointwise kernel between them.
"""
import torch
import torch.nn as nn

import ops
from _heads import LinearF32


class _LstmInput(torch.autograd.Function):
    """cat([action.float() / 48, flatten(new_tensor)], dim=1)  (rovr/action_lstm.py:28-31)."""

    @staticmethod
    def forward(ctx, action, new_tensor):
        b = action.shape[0]
        a2 = action.float().contiguous()
        t2 = new_tensor.float().contiguous().view(b, -1)
        out = torch.empty((b, a2.shape[1] + t2.shape[1]), dtype=torch.float32, device=t2.device)
        ops.copy2d_f32(a2, out[:, : a2.shape[1]], scale=1.0 / 48)
        ops.copy2d_f32(t2, out[:, a2.shape[1]:])
        ctx.na, ctx.tshape = a2.shape[1], new_tensor.shape
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        gt = torch.empty((g.shape[0], g.shape[1] - ctx.na), dtype=torch.float32, device=g.device)
        ops.copy2d_f32(g[:, ctx.na:], gt)
        return None, gt.view(ctx.tshape)


class _LstmCell(torch.autograd.Function):
    """(gates [b,4H] pre-activation, c_prev [b,H]) -> (h, c); gate order i, f, g, o."""

    @staticmethod
    def forward(ctx, gates, c_prev):
        gates, c_prev = gates.contiguous(), c_prev.contiguous()
        h, c, act = ops.lstm_pointwise_fwd(gates, c_prev)
        ctx.save_for_backward(act, c_prev, c)
        return h, c

    @staticmethod
    def backward(ctx, dh, dc):
        act, c_prev, c = ctx.saved_tensors
        dh = dh.contiguous() if dh is not None else None
        dc = dc.contiguous() if dc is not None else None
        dgates, dc_prev = ops.lstm_pointwise_bwd(act, c_prev, c, dh, dc)
        return dgates, dc_prev


class _AddRows(torch.autograd.Function):
    """a + b for two fp32 [b, n] tensors (the two halves of the LSTMCell gate pre-activation)."""

    @staticmethod
    def forward(ctx, a, b):
        out = torch.empty_like(a)
        ops.copy2d_f32(a.contiguous(), out)
        ops.copy2d_f32(b.contiguous(), out, accumulate=True)
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


class ActionLSTM(nn.Module):
    """Reference: rovr/action_lstm.py:6-42."""

    def __init__(self, hidden_dim, num_layers, batch_size):
        super(ActionLSTM, self).__init__()
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        self.batch_size = batch_size
        self.lstm = nn.LSTMCell(3 + 3 * 16 * 16 * 3, hidden_dim)
        self.fc = nn.Linear(hidden_dim, 80 * 80 * 3)
        self.hx = torch.zeros(batch_size, self.hidden_dim)
        self.cx = torch.zeros(batch_size, self.hidden_dim)

    def forward(self, action, new_tensor):
        if not action.is_cuda:
            raise RuntimeError("ActionLSTM (B200) needs CUDA tensors: there is no CPU path")
        local_device = action.device
        self.hx = self.hx.to(local_device)
        self.cx = self.cx.to(local_device)
        input_tensor = _LstmInput.apply(action, new_tensor)
        gates = _AddRows.apply(LinearF32.apply(input_tensor, self.lstm.weight_ih, self.lstm.bias_ih),
                               LinearF32.apply(self.hx, self.lstm.weight_hh, self.lstm.bias_hh))
        self.hx, self.cx = _LstmCell.apply(gates, self.cx)
        out = LinearF32.apply(self.hx, self.fc.weight, self.fc.bias)
        return out.view(-1, 3, 80, 80)                               # 'b (c ph pw) -> b c ph pw'

    def reset_hidden_states(self):
        self.hx = torch.zeros(self.batch_size, self.hidden_dim)
        self.cx = torch.zeros(self.batch_size, self.hidden_dim)
