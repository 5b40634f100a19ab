"""``ConvDiscriminator`` used by the training step (reference model/multimodal_context_net.py:219-268).

Not on the hot path (stays stock PyTorch); mirrored here only so the step in
hop_b200.train_llm is self-contained.  Same modules, names and initialisation order as the reference.
"""
import torch
import torch.nn as nn


class ConvDiscriminator(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = 64
        self.pre_conv = nn.Sequential(
            nn.Conv1d(input_size, 16, 3), nn.BatchNorm1d(16), nn.LeakyReLU(True),
            nn.Conv1d(16, 8, 3), nn.BatchNorm1d(8), nn.LeakyReLU(True),
            nn.Conv1d(8, 8, 3))
        self.gru = nn.GRU(8, hidden_size=self.hidden_size, num_layers=4, bidirectional=True, dropout=0.3, batch_first=True)
        self.out = nn.Linear(self.hidden_size, 1)
        self.out2 = nn.Linear(28, 1)

    def forward(self, poses, in_text=None):
        feat = self.pre_conv(poses.transpose(1, 2)).transpose(1, 2)
        output, _ = self.gru(feat, None)
        output = output[:, :, :self.hidden_size] + output[:, :, self.hidden_size:]
        batch_size = poses.shape[0]
        output = self.out(output.contiguous().view(-1, output.shape[2])).view(batch_size, -1)
        return torch.sigmoid(self.out2(output))
