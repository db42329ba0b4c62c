"""CNNFeatureExtractor — the depth-image encoder of the avoid/planning policies, with the reference's module layout so that
its checkpoints load key for key (lib/network/cnn.py:3-33: `features.{0,3,6}` convolutions 1→16 (5x5, s2) →32 (3x3, s2) →64
(3x3, s2), each followed by ReLU then BatchNorm2d `features.{2,5,8}`, global average pool, `fc` 64→feature_dim).

The convolutions here are cuDNN's (library calls through torch): SURVEY.md §8(f) row 3 — hand-written tensor-core kernels for
this encoder are not built yet; the MLP trunk behind it does run on the libagx kernels."""
import torch.nn as nn


class CNNFeatureExtractor(nn.Module):
    def __init__(self, feature_dim=12):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(1, 16, kernel_size=5, stride=2, padding=2), nn.ReLU(), nn.BatchNorm2d(16),   # (16, 106, 60)
            nn.Conv2d(16, 32, kernel_size=3, stride=2, padding=1), nn.ReLU(), nn.BatchNorm2d(32),  # (32, 53, 30)
            nn.Conv2d(32, 64, kernel_size=3, stride=2, padding=1), nn.ReLU(), nn.BatchNorm2d(64),  # (64, 27, 15)
            nn.AdaptiveAvgPool2d((1, 1)),
        )
        self.fc = nn.Linear(64, feature_dim)

    def forward(self, x):
        x = self.features(x)
        return self.fc(x.view(x.size(0), -1))
