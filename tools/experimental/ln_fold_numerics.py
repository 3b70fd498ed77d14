"""CPU numerics study for the LayerNorm-folding plan: does running the K = 512 consumer GEMMs on bf16(x) with
gamma folded into the weights and (mean, rstd) applied in the epilogue stay as close to the fp32 reference as today's
bf16(LayerNorm(x)) operand?  Emulates bf16 operands / fp32 accumulation with torch on the CPU (oracle/model.py structure).

    python tools/experimental/ln_fold_numerics.py            # prints max |probs - fp32|, max |bounds - fp32| per scheme
"""
import pathlib
import sys

import torch
import torch.nn.functional as F

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import model as om  # noqa: E402
from some_b200 import synth  # noqa: E402


def bf(x):
    return x.bfloat16().float()


class Emu:
    """mode 'fp32' | 'bf16' (today: a = bf16(LN(x))) | 'fold' (bf16(x), gamma in W, stats in the epilogue)."""

    def __init__(self, sd, mode, mean_shift=0.0):
        self.sd, self.mode, self.mean_shift = sd, mode, mean_shift

    def lin(self, name, x, bias=True):
        w, b = self.sd[name + '.weight'], (self.sd[name + '.bias'] if bias else None)
        if self.mode == 'fp32':
            return F.linear(x, w, b)
        return F.linear(bf(x), bf(w), b)

    def ln_lin(self, ln, names, x, bias=True):
        """LayerNorm(ln) followed by one or more Linear layers on the normalised rows."""
        g, be = self.sd[ln + '.weight'], self.sd[ln + '.bias']
        if self.mode != 'fold':
            a = F.layer_norm(x, (x.shape[-1],), g, be, 1e-5)
            return [self.lin(n, a, bias) for n in names]
        mu = x.mean(-1, keepdim=True)
        var = (x * x).mean(-1, keepdim=True) - mu * mu                  # one-pass variance from (sum x, sum x^2)
        rstd = torch.rsqrt(var + 1e-5)
        xb = bf(x)
        outs = []
        for n in names:
            w = self.sd[n + '.weight']
            wf = bf(w * g)                                              # W' = W * gamma, rounded once offline
            s = wf.sum(-1)                                              # s_n = sum_k W'_nk (fp32)
            b2 = w @ be + (self.sd[n + '.bias'] if bias else 0.0)       # b'_n = bias_n + sum_k beta_k W_nk
            acc = F.linear(xb, wf)
            outs.append(rstd * (acc - mu * s) + b2)
        return outs

    def block(self, p, x, heads):
        sd = self.sd
        (h,) = self.ln_lin(p + '.norm1', [p + '.ffn1.ln1'], x)
        x = self.lin(p + '.ffn1.ln2', F.silu(h)) * 0.5 + x
        q, kv = self.ln_lin(p + '.norm2', [p + '.att.to_q', p + '.att.to_kv'], x, bias=False)
        if self.mode != 'fp32':
            q, kv = bf(q), bf(kv)
        k, v = kv.chunk(2, dim=2)
        b, t, _ = x.shape
        q, k, v = (z.reshape(b, t, heads, -1).transpose(1, 2) for z in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, t, -1)
        x = self.lin(p + '.att.to_out.0', o) + x
        w1, b1 = sd[p + '.conv.pointwise_conv1.weight'].squeeze(-1), sd[p + '.conv.pointwise_conv1.bias']
        self.sd[p + '.conv.pw1.weight'], self.sd[p + '.conv.pw1.bias'] = w1, b1
        (c,) = self.ln_lin(p + '.norm3', [p + '.conv.pw1'], x)
        c = om._glu(c, 2).transpose(1, 2)
        kk = sd[p + '.conv.depthwise_conv.weight'].shape[-1]
        c = F.conv1d(bf(c) if self.mode != 'fp32' else c, sd[p + '.conv.depthwise_conv.weight'], sd[p + '.conv.depthwise_conv.bias'],
                     padding=(kk - 1) // 2, groups=c.shape[1])
        c = F.batch_norm(c, sd[p + '.conv.norm.running_mean'], sd[p + '.conv.norm.running_var'], sd[p + '.conv.norm.weight'],
                         sd[p + '.conv.norm.bias'], False, 0.1, 1e-5)
        c = F.silu(c).transpose(1, 2)
        self.sd[p + '.conv.pw2.weight'] = sd[p + '.conv.pointwise_conv2.weight'].squeeze(-1)
        self.sd[p + '.conv.pw2.bias'] = sd[p + '.conv.pointwise_conv2.bias']
        x = self.lin(p + '.conv.pw2', c) + x
        (h,) = self.ln_lin(p + '.norm4', [p + '.ffn2.ln1'], x)
        x = self.lin(p + '.ffn2.ln2', F.silu(h)) * 0.5 + x
        x = F.layer_norm(x, (x.shape[-1],), sd[p + '.norm5.weight'], sd[p + '.norm5.bias'], 1e-5)
        return x + self.mean_shift                                       # stress: rows with a large common offset

    def forward(self, units, lay, heads):
        x, x1 = self.lin('model.inln', units), self.lin('model.inln1', units)
        for i in range(lay):
            p = f'model.cf_lay.{i}'
            midi, bound = self.block(p + '.att1', x, heads), self.block(p + '.att2', x1, heads)
            x = midi + om._glu(self.lin(p + '.glu2.0', bound), 2)
            x1 = bound + om._glu(self.lin(p + '.glu1.0', midi), 2)
        x, x1 = self.block('model.att1', x, heads), self.block('model.att2', x1, heads)
        return torch.sigmoid(self.lin('model.outln', x)), torch.sigmoid(self.lin('model.cutheard', x1)).squeeze(-1)


def main():
    torch.manual_seed(0)
    for name in ('two_head', 'midi_conformer'):
        config = synth.named_config(name)
        sd = {k: v.float() for k, v in synth.fabricate_state_dict(config).items()}
        wave = torch.from_numpy(synth.synth_waveform(77, seconds=6.0))[None]
        units = om.log_mel(wave).transpose(1, 2)
        lay, heads = config['midi_extractor_args']['lay'], config['midi_extractor_args']['attention_heads']
        for shift in (0.0, 4.0):
            ref = Emu(dict(sd), 'fp32', shift).forward(units, lay, heads)
            line = f'{name:15s} row offset {shift:3.1f}:'
            for mode in ('bf16', 'fold'):
                got = Emu(dict(sd), mode, shift).forward(units, lay, heads)
                line += f'  {mode}: probs {float((got[0] - ref[0]).abs().max()):.2e} bounds {float((got[1] - ref[1]).abs().max()):.2e}'
            print(line)


if __name__ == '__main__':
    with torch.no_grad():
        main()
