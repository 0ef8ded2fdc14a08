"""Oracle: the consumer of the features, restated functionally (TEST INFRASTRUCTURE ONLY).

`/root/reference/cnn_bilstm_hybrid.py:20-68` (`AudioDeepfakeDetector`): Conv1d over the 63 time frames
as channels -> BatchNorm -> ReLU -> MaxPool(2) -> BiLSTM(64 -> 2 x 32) -> linear attention ->
softmax over time -> LayerNorm(1) -> weighting -> max over time -> MLP -> sigmoid, in eval mode
(dropout off, BatchNorm on running statistics).  Weights come from a state dict (the committed fixture
`tests/golden/consumer.npz`, produced by the real reference class); this file only re-expresses the
forward pass with torch.nn.functional so that the GPU box, which has no /root/reference, can check
that features from the CUDA path drive the model to the same scores.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def forward(weights: dict, x: torch.Tensor) -> torch.Tensor:
    """x (B, F, 63) float32 -> scores (B, 1).  `weights`: name -> tensor as in the reference state dict."""
    w = {k: v.to(device=x.device, dtype=torch.float32) for k, v in weights.items()}
    h = x.permute(0, 2, 1)                                                   # (B, 63, F)   :56
    h = F.conv1d(h, w["feature_extractor.0.weight"], w["feature_extractor.0.bias"], padding=1)
    h = F.batch_norm(h, w["feature_extractor.1.running_mean"], w["feature_extractor.1.running_var"],
                     w["feature_extractor.1.weight"], w["feature_extractor.1.bias"], training=False)
    h = F.max_pool1d(F.relu(h), kernel_size=2)                               # (B, 64, F//2)
    h = h.permute(0, 2, 1)                                                   # (B, F//2, 64) :58
    hidden = w["bilstm.weight_hh_l0"].shape[1]
    lstm = torch.nn.LSTM(64, hidden, num_layers=1, batch_first=True, bidirectional=True).to(x.device)
    with torch.no_grad():
        for name, p in lstm.named_parameters():
            p.copy_(w["bilstm." + name])
    lstm.eval()
    out, _ = lstm(h)                                                         # (B, F//2, 2*hidden) :60
    attn = torch.softmax(F.linear(out, w["attention.weight"], w["attention.bias"]), dim=1)
    attn = F.layer_norm(attn, (1,), w["layer_norm.weight"], w["layer_norm.bias"])
    pooled = torch.max(out * attn, dim=1).values                             # :66
    z = F.relu(F.linear(pooled, w["classifier.0.weight"], w["classifier.0.bias"]))
    return torch.sigmoid(F.linear(z, w["classifier.3.weight"], w["classifier.3.bias"]))
