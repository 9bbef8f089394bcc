"""LanguageModelCriterion -- drop-in for lib/misc/utils.py:39-53 (masked NLL over caption log-probs)."""
import torch
import torch.nn as nn


class LanguageModelCriterion(nn.Module):
    def forward(self, input, target, mask):
        T = input.size(1)
        target = target[:, :T]
        mask = mask[:, :T].to(input.dtype)
        picked = input.gather(2, target.unsqueeze(2)).squeeze(2)
        return -(picked * mask).sum() / mask.sum()
