"""BERT for pre-training (masked LM + next-sentence heads), BERT-large by default: 24 layers, hidden 1024, 16 heads,
FFN 4096, vocab 30522, max 512 positions, ~336 M parameters — the `BERT-large pretraining bf16` config of
BASELINE.json.  Attention goes through torch's fused scaled_dot_product_attention."""
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden_size: int = 1024
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    intermediate_size: int = 4096
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    layer_norm_eps: float = 1e-12


class BertEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids):
        pos = torch.arange(input_ids.shape[1], device=input_ids.device).unsqueeze(0)
        x = self.word_embeddings(input_ids) + self.position_embeddings(pos) + self.token_type_embeddings(token_type_ids)
        return self.dropout(self.LayerNorm(x))


class BertLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.nh = c.num_attention_heads
        self.qkv = nn.Linear(c.hidden_size, 3 * c.hidden_size)
        self.attn_out = nn.Linear(c.hidden_size, c.hidden_size)
        self.attn_ln = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.ffn_in = nn.Linear(c.hidden_size, c.intermediate_size)
        self.ffn_out = nn.Linear(c.intermediate_size, c.hidden_size)
        self.ffn_ln = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.drop = nn.Dropout(c.hidden_dropout_prob)
        self.attn_p = c.attention_probs_dropout_prob

    def forward(self, x, mask):
        b, s, h = x.shape
        q, k, v = self.qkv(x).view(b, s, 3, self.nh, h // self.nh).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=self.attn_p if self.training else 0.0)
        a = a.transpose(1, 2).reshape(b, s, h)
        x = self.attn_ln(x + self.drop(self.attn_out(a)))
        return self.ffn_ln(x + self.drop(self.ffn_out(F.gelu(self.ffn_in(x)))))


class BertModel(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.embeddings = BertEmbeddings(c)
        self.layers = nn.ModuleList([BertLayer(c) for _ in range(c.num_hidden_layers)])
        self.pooler = nn.Linear(c.hidden_size, c.hidden_size)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None):
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        mask = None
        if attention_mask is not None:
            mask = attention_mask[:, None, None, :].to(torch.bool)
        x = self.embeddings(input_ids, token_type_ids)
        for layer in self.layers:
            x = layer(x, mask)
        return x, torch.tanh(self.pooler(x[:, 0]))


class BertForPreTraining(nn.Module):
    def __init__(self, c=None):
        super().__init__()
        c = c or BertConfig()
        self.config = c
        self.bert = BertModel(c)
        self.transform = nn.Linear(c.hidden_size, c.hidden_size)
        self.transform_ln = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.decoder_bias = nn.Parameter(torch.zeros(c.vocab_size))
        self.seq_relationship = nn.Linear(c.hidden_size, 2)
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, std=0.02)
        if isinstance(m, nn.Linear) and m.bias is not None:
            nn.init.zeros_(m.bias)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, labels=None, next_sentence_label=None):
        seq, pooled = self.bert(input_ids, token_type_ids, attention_mask)
        h = self.transform_ln(F.gelu(self.transform(seq)))
        logits = F.linear(h, self.bert.embeddings.word_embeddings.weight, self.decoder_bias)  # tied decoder
        nsp = self.seq_relationship(pooled)
        if labels is None:
            return logits, nsp
        loss = F.cross_entropy(logits.view(-1, logits.shape[-1]).float(), labels.view(-1), ignore_index=-100)
        if next_sentence_label is not None:
            loss = loss + F.cross_entropy(nsp.float(), next_sentence_label.view(-1))
        return loss


def bert_large():
    return BertForPreTraining(BertConfig())
