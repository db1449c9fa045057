"""GPT-2 (decoder-only LM), medium by default: 24 layers, width 1024, 16 heads, context 1024, vocab 50257,
~355 M parameters — the `GPT-2 medium with Adasum` config of BASELINE.json."""
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class GPT2Config:
    vocab_size: int = 50257
    n_positions: int = 1024
    n_embd: int = 1024
    n_layer: int = 24
    n_head: int = 16
    dropout: float = 0.1


class Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.nh = c.n_head
        self.ln_1 = nn.LayerNorm(c.n_embd)
        self.c_attn = nn.Linear(c.n_embd, 3 * c.n_embd)
        self.c_proj = nn.Linear(c.n_embd, c.n_embd)
        self.ln_2 = nn.LayerNorm(c.n_embd)
        self.c_fc = nn.Linear(c.n_embd, 4 * c.n_embd)
        self.c_proj2 = nn.Linear(4 * c.n_embd, c.n_embd)
        self.drop = nn.Dropout(c.dropout)
        self.p = c.dropout

    def forward(self, x):
        b, s, h = x.shape
        q, k, v = self.c_attn(self.ln_1(x)).view(b, s, 3, self.nh, h // self.nh).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True, dropout_p=self.p if self.training else 0.0)
        x = x + self.drop(self.c_proj(a.transpose(1, 2).reshape(b, s, h)))
        return x + self.drop(self.c_proj2(F.gelu(self.c_fc(self.ln_2(x)), approximate='tanh')))


class GPT2LMHeadModel(nn.Module):
    def __init__(self, c=None):
        super().__init__()
        c = c or GPT2Config()
        self.config = c
        self.wte = nn.Embedding(c.vocab_size, c.n_embd)
        self.wpe = nn.Embedding(c.n_positions, c.n_embd)
        self.drop = nn.Dropout(c.dropout)
        self.h = nn.ModuleList([Block(c) for _ in range(c.n_layer)])
        self.ln_f = nn.LayerNorm(c.n_embd)
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, std=0.02)
        if isinstance(m, nn.Linear) and m.bias is not None:
            nn.init.zeros_(m.bias)

    def forward(self, input_ids, labels=None):
        pos = torch.arange(input_ids.shape[1], device=input_ids.device).unsqueeze(0)
        x = self.drop(self.wte(input_ids) + self.wpe(pos))
        for blk in self.h:
            x = blk(x)
        logits = F.linear(self.ln_f(x), self.wte.weight)  # tied LM head
        if labels is None:
            return logits
        return F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), labels[:, 1:].reshape(-1))


def gpt2_medium():
    return GPT2LMHeadModel(GPT2Config())
