"""Call patterns of the reference's GENERATIVE stack around the drop-ins: the stateful decoder
(decoders.py:28-76), the model-side step / beam-search entry (base_transformer.py:30-54) and the beam-search driver
(beam_search.py:41-118).

Scaffolding for the GPU parity tests and the decode microbenchmark — a reference checkout keeps its own ``Decoder`` /
``BaseTransformer`` / ``BeamSearch`` and gets the fast path through ``install()`` (``MultiHeadAttention`` /
``PositionWiseFeedForward`` drop-ins, ``BeamSearch.select`` on the top-k kernel, tracked state re-gathers).  Same parameter names as
the reference, so reference-generated golden vectors load by name.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from ..modules import MASK_VALUE, StatefulModule, beam_select, tracked_regather
from .models import DecoderLayer, generate_padding_mask, generate_sequential_mask


class StatefulModuleList(nn.ModuleList, StatefulModule):
    """containers.py:73-74: a ModuleList the state recursion walks through."""


def sinusoid_encoding_table(max_len: int, d_model: int, padding_idx=None) -> torch.Tensor:
    """models/utils.py:21-38: interleaved sin / cos of pos / 10000^(2k/d), row ``padding_idx`` zeroed."""
    pos = torch.arange(max_len, dtype=torch.float32).view(-1, 1)
    k = torch.arange(d_model // 2, dtype=torch.float32).view(1, -1)
    table = torch.zeros((max_len, d_model))
    table[:, 0::2] = torch.sin(pos / 10000 ** (2 * k / d_model))
    table[:, 1::2] = torch.cos(pos / 10000 ** (2 * k / d_model))
    if padding_idx is not None:
        table[padding_idx] = 0
    return table


class UsualEmbedding(nn.Module):
    """text_embeddings.py:57-80 without pretrained vectors: a lookup table under the name ``components``."""

    def __init__(self, d_model, vocab):
        super().__init__()
        self.components = nn.Embedding(len(vocab), d_model, vocab.padding_idx)

    def forward(self, tokens):
        return self.components(tokens)


class Decoder(StatefulModule):
    """decoders.py:28-76.  Inside ``statefulness()`` it is fed one token per call: the self-attention mask and the position counter
    are running states (``running_mask_self_attention``, ``running_seq``), the layers' attention blocks keep theirs."""

    def __init__(self, config, vocab):
        super().__init__()
        self.d_model, self.max_len, self.padding_idx = config.D_MODEL, vocab.max_answer_length, vocab.padding_idx
        self.word_emb = UsualEmbedding(config.D_MODEL, vocab)
        self.pos_emb = nn.Embedding.from_pretrained(sinusoid_encoding_table(self.max_len + 1, config.D_MODEL, 0), freeze=True)
        self.layers = StatefulModuleList([DecoderLayer(config.ATTENTION) for _ in range(config.LAYERS)])
        self.fc = nn.Linear(config.D_MODEL, len(vocab), bias=False)
        self.register_state("running_mask_self_attention", torch.zeros((1, 1, 0)).bool())
        self.register_state("running_seq", torch.zeros((1,)).long())

    def forward(self, answer_tokens, encoder_features, encoder_attention_mask):
        B, L = answer_tokens.shape
        dev = answer_tokens.device
        pad = generate_padding_mask(answer_tokens, self.padding_idx).to(dev)                      # (B,1,1,L)
        self_mask = torch.logical_or(pad != 0, generate_sequential_mask(L).to(dev) != 0).long() * MASK_VALUE   # models/utils.py:68-73
        if self._is_stateful:
            self.running_mask_self_attention = torch.cat([self.running_mask_self_attention, self_mask], -1)
            self_mask = self.running_mask_self_attention
        positions = torch.arange(1, L + 1, device=dev).view(1, -1).expand(B, -1).masked_fill(pad.view(B, L) != 0, 0)
        if self._is_stateful:
            self.running_seq.add_(1)
            positions = self.running_seq
        x = self.word_emb(answer_tokens) + self.pos_emb(positions)
        for layer in self.layers:
            x = layer(x, encoder_features, encoder_features, self_mask, encoder_attention_mask)
        return F.log_softmax(self.fc(x), dim=-1)


class Captioner(StatefulModule):
    """base_transformer.py:8-54 around a given encoder output: ``encoder_features`` / ``encoder_padding_mask`` are states of the
    model (so beam search re-gathers them), ``step`` feeds the previous words to the decoder."""

    def __init__(self, config, vocab):
        super().__init__()
        self.vocab, self.max_len, self.eos_idx = vocab, vocab.max_answer_length, vocab.eos_idx
        self.register_state("encoder_features", None)
        self.register_state("encoder_padding_mask", None)
        self.decoder = Decoder(config, vocab)

    def step(self, t, prev_output):
        if t == 0:
            prev_output = torch.full((self.encoder_features.shape[0], 1), self.vocab.bos_idx, dtype=torch.long,
                                     device=self.encoder_features.device)
        return self.decoder(prev_output, self.encoder_features, self.encoder_padding_mask)

    def beam_search(self, encoder_features, encoder_padding_mask, batch_size, beam_size, out_size=1):
        search = BeamSearch(self, batch_size, self.max_len, self.eos_idx, beam_size, encoder_features.device)
        with self.statefulness(batch_size):
            self.encoder_features, self.encoder_padding_mask = encoder_features, encoder_padding_mask
            return search.apply(out_size)


class BeamSearch:
    """beam_search.py:4-118.  Per step: score every (beam, word) continuation, freeze finished beams, keep the ``beam_size`` best
    per sample (``beam_select``: the row-wise top-k kernel instead of a full sort), re-gather every state of the model to the
    surviving beams with ONE tracked closure, and carry the chosen words / log-probabilities along."""

    def __init__(self, model, b_s, max_len, eos_idx, beam_size, device):
        self.model, self.b_s, self.max_len, self.eos_idx, self.beam_size, self.device = model, b_s, max_len, eos_idx, beam_size, device

    def _regather(self, beam_of, cur_beams):
        def fn(state):
            tail = list(state.shape[1:])
            index = beam_of.view(self.b_s, self.beam_size, *([1] * len(tail))).expand(self.b_s, self.beam_size, *tail)
            return torch.gather(state.view(self.b_s, cur_beams, *tail), 1, index).view(-1, *tail)     # (as the reference: no empty states)
        return tracked_regather(fn)

    def select(self, candidate_logprob):
        return beam_select(candidate_logprob, self.b_s, self.beam_size)

    def _step(self, t, words_so_far):
        cur = 1 if t == 0 else self.beam_size
        word_lp = self.model.step(t, self.selected_words).view(self.b_s, cur, -1)
        cand = self.seq_logprob + word_lp
        if t > 0:                                                          # a beam that has emitted <eos> keeps its score: one candidate
            alive = (self.selected_words.view(self.b_s, cur) != self.eos_idx).float().unsqueeze(-1)
            self.seq_mask = self.seq_mask * alive
            word_lp = word_lp * self.seq_mask.expand_as(word_lp)
            frozen = self.seq_logprob.expand_as(cand).contiguous()
            frozen[:, :, 1:] = -999
            cand = self.seq_mask * cand + frozen * (1 - self.seq_mask)
        flat_idx, lp = self.select(cand)
        V = cand.shape[-1]
        beam_of = torch.div(flat_idx, V, rounding_mode="trunc")
        word_of = flat_idx - beam_of * V
        self.model.apply_to_states(self._regather(beam_of, cur))
        pick = beam_of.unsqueeze(-1)
        self.seq_logprob = lp.unsqueeze(-1)
        self.seq_mask = torch.gather(self.seq_mask, 1, pick)
        words_so_far = [torch.gather(w, 1, pick) for w in words_so_far] + [word_of.unsqueeze(-1)]
        step_lp = torch.gather(torch.gather(word_lp, 1, pick.expand(self.b_s, self.beam_size, V)), 2, word_of.unsqueeze(-1))
        self.log_probs = [torch.gather(p, 1, pick.expand(self.b_s, self.beam_size, 1)) for p in self.log_probs] + [step_lp]
        self.selected_words = word_of.view(-1, 1)
        return words_so_far

    def apply(self, out_size=1):
        self.seq_mask = torch.ones((self.b_s, self.beam_size, 1), device=self.device)
        self.seq_logprob = torch.zeros((self.b_s, 1, 1), device=self.device)
        self.log_probs, self.selected_words, words = [], None, []
        for t in range(self.max_len):
            words = self._step(t, words)
        _, order = torch.sort(self.seq_logprob, 1, descending=True)
        order = order.expand(self.b_s, self.beam_size, self.max_len)
        tokens = torch.gather(torch.cat(words, -1), 1, order).contiguous()[:, :out_size]
        log_probs = torch.gather(torch.cat(self.log_probs, -1), 1, order).contiguous()[:, :out_size]
        if out_size == 1:
            tokens, log_probs = tokens.squeeze(1), log_probs.squeeze(1)
        return tokens, log_probs
