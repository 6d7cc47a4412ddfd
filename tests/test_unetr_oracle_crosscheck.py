"""oracle/unetr_oracle.py is a restatement of MONAI blocks that cannot be pinned on MONAI itself (absent from the reference
tree and from this image).  What CAN be pinned: every block of it against the stock torch.nn module that implements the
same published definition, with mapped weights --
  * TransformerBlock (pre-norm, SABlock without qkv bias, MLPBlock with GELU)  ==  nn.TransformerEncoderLayer(norm_first=True,
    activation="gelu") whose nn.MultiheadAttention has in_proj_weight = qkv.weight, in_proj_bias = 0;
  * UnetResBlock  ==  nn.Conv3d(bias=False) / nn.InstanceNorm3d(affine=False) / nn.LeakyReLU(0.01) wired as published;
  * UnetrPrUpBlock / UnetrUpBlock transposed convolutions  ==  nn.ConvTranspose3d(kernel 2, stride 2, bias=False);
  * the perceptron patch embedding  ==  einops' 'b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)' + nn.Linear.
The network-level wiring (which hidden states are tapped, skip order) follows code/networks/unetr.py:215-230 and is what
tests/test_unetr_gpu.py checks the CUDA path against."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import unetr_oracle as UO


def test_vit_blocks_match_torch_transformer_encoder_layer():
    torch.manual_seed(0)
    hid, heads, mlp, layers, B, P, S = 48, 4, 96, 3, 2, 16, 32
    ntok = (S // P) ** 3
    sd = {"vit.patch_embedding.patch_embeddings.1.weight": torch.randn(hid, P ** 3) * 0.02,
          "vit.patch_embedding.patch_embeddings.1.bias": torch.randn(hid) * 0.02,
          "vit.patch_embedding.position_embeddings": torch.randn(1, ntok, hid) * 0.02,
          "vit.norm.weight": torch.rand(hid) + 0.5, "vit.norm.bias": torch.randn(hid) * 0.1}
    enc = []
    for i in range(layers):
        layer = nn.TransformerEncoderLayer(hid, heads, mlp, dropout=0.0, activation="gelu", batch_first=True, norm_first=True)
        layer.self_attn.in_proj_bias.data.zero_()                    # MONAI's SABlock: qkv Linear without bias
        p = f"vit.blocks.{i}."
        sd[p + "norm1.weight"], sd[p + "norm1.bias"] = layer.norm1.weight.data, layer.norm1.bias.data
        sd[p + "norm2.weight"], sd[p + "norm2.bias"] = layer.norm2.weight.data, layer.norm2.bias.data
        sd[p + "attn.qkv.weight"] = layer.self_attn.in_proj_weight.data      # rows (qkv, head, dim): '(qkv l d)'
        sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"] = layer.self_attn.out_proj.weight.data, layer.self_attn.out_proj.bias.data
        sd[p + "mlp.linear1.weight"], sd[p + "mlp.linear1.bias"] = layer.linear1.weight.data, layer.linear1.bias.data
        sd[p + "mlp.linear2.weight"], sd[p + "mlp.linear2.bias"] = layer.linear2.weight.data, layer.linear2.bias.data
        for prm in (layer.norm1.weight, layer.norm1.bias, layer.norm2.weight, layer.norm2.bias):
            prm.data.add_(torch.randn_like(prm) * 0.1)
        enc.append(layer.eval())
    x = torch.randn(B, 1, S, S, S)
    out, hidden = UO.vit_forward(sd, x, heads, num_layers=layers, patch=P)
    # reference wiring with stock modules
    from einops import rearrange
    t = rearrange(x, "b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)", p1=P, p2=P, p3=P)
    t = F.linear(t, sd["vit.patch_embedding.patch_embeddings.1.weight"], sd["vit.patch_embedding.patch_embeddings.1.bias"])
    t = t + sd["vit.patch_embedding.position_embeddings"]
    with torch.no_grad():
        for i, layer in enumerate(enc):
            t = layer(t)
            torch.testing.assert_close(hidden[i], t, rtol=1e-4, atol=1e-5)
        t = F.layer_norm(t, (hid,), sd["vit.norm.weight"], sd["vit.norm.bias"], 1e-5)
    torch.testing.assert_close(out, t, rtol=1e-4, atol=1e-5)


class _Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv3d(cin, cout, 3, 1, 1, bias=False), nn.Conv3d(cout, cout, 3, 1, 1, bias=False)
        self.conv3 = nn.Conv3d(cin, cout, 1, 1, 0, bias=False) if cin != cout else None
        self.norm1, self.norm2, self.norm3 = (nn.InstanceNorm3d(cout, affine=False) for _ in range(3))
        self.act = nn.LeakyReLU(0.01)

    def forward(self, x):
        out = self.act(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        res = self.norm3(self.conv3(x)) if self.conv3 is not None else x
        return self.act(out + res)


def _res_sd(sd, prefix, m):
    sd[prefix + "conv1.conv.weight"], sd[prefix + "conv2.conv.weight"] = m.conv1.weight.data, m.conv2.weight.data
    if m.conv3 is not None:
        sd[prefix + "conv3.conv.weight"] = m.conv3.weight.data


def test_conv_blocks_match_torch_modules():
    torch.manual_seed(1)
    x = torch.randn(2, 6, 8, 8, 8)
    for cin, cout in ((6, 6), (6, 10)):
        m, sd = _Res(cin, cout).eval(), {}
        _res_sd(sd, "b.", m)
        with torch.no_grad():
            torch.testing.assert_close(UO.res_block(sd, "b.", x), m(x), rtol=1e-4, atol=1e-5)
    # UnetrPrUpBlock(num_layer = 2): transp_conv_init, then 2 x [ConvTranspose3d k2 s2, UnetResBlock]
    init, ups, ress = nn.ConvTranspose3d(6, 4, 2, 2, bias=False), [nn.ConvTranspose3d(4, 4, 2, 2, bias=False) for _ in range(2)], [_Res(4, 4) for _ in range(2)]
    sd = {"e.transp_conv_init.conv.weight": init.weight.data}
    for j in range(2):
        sd[f"e.blocks.{j}.0.conv.weight"] = ups[j].weight.data
        _res_sd(sd, f"e.blocks.{j}.1.", ress[j])
    xs = torch.randn(1, 6, 3, 3, 3)
    with torch.no_grad():
        ref = init(xs)
        for j in range(2):
            ref = ress[j](ups[j](ref))
        torch.testing.assert_close(UO.pr_up_block(sd, "e.", xs, 2), ref, rtol=1e-4, atol=1e-5)
    # UnetrUpBlock: ConvTranspose3d, cat((out, skip), 1), UnetResBlock(2 * cout -> cout)
    up, res = nn.ConvTranspose3d(6, 4, 2, 2, bias=False), _Res(8, 4)
    sd = {"d.transp_conv.conv.weight": up.weight.data}
    _res_sd(sd, "d.conv_block.", res)
    skip = torch.randn(1, 4, 6, 6, 6)
    with torch.no_grad():
        torch.testing.assert_close(UO.up_block(sd, "d.", xs, skip), res(torch.cat((up(xs), skip), 1)), rtol=1e-4, atol=1e-5)
