"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the reference's UNETR path,
code/networks/unetr.py:215-230 built by code/networks/net_factory_3d.py:27-39.

PARITY UNPINNED: the arithmetic lives in MONAI (`monai.networks.nets.ViT`, `monai.networks.blocks.{UnetrBasicBlock,
UnetrPrUpBlock, UnetrUpBlock}`, `dynunet_block.{UnetResBlock, UnetOutBlock}`), which is neither vendored under
/root/reference nor installed here, and the reference pins no MONAI version (no requirements file; the `pos_embed=`
keyword limits it to roughly 0.6 - 1.3).  The reference holds no test or golden vector for this network.  The functions
below restate MONAI's published block definitions (cited per function) and are anchored on the reference's own call
site (unetr.py:88-181 constructor arguments, :215-230 forward); state_dict keys follow MONAI's module names.
What is pinned instead (tests/test_unetr_oracle_crosscheck.py): every block below against the stock torch.nn module
of the same published definition with mapped weights -- the transformer blocks against nn.TransformerEncoderLayer
(norm_first, gelu, zero in_proj bias), the residual / transposed-convolution blocks against nn.Conv3d,
nn.InstanceNorm3d, nn.ConvTranspose3d.  That removes "the restatement mis-implements a block" but not "MONAI's
version differs from the published definition", so the header stays "unpinned".
"""
import torch
import torch.nn.functional as F

LRELU = 0.01


def vit_forward(sd, x, heads, num_layers=12, patch=16):
    """monai ViT.forward with classification=False: PatchEmbeddingBlock('perceptron') -> 12 TransformerBlocks -> LayerNorm.
    PatchEmbeddingBlock: Rearrange 'b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)' + Linear, + position_embeddings.
    TransformerBlock: x = x + attn(norm1(x)); x = x + mlp(norm2(x)).  SABlock: qkv Linear without bias, columns
    '(qkv l d)', softmax(q k^T * d^-0.5) v, out_proj.  MLPBlock: linear1 -> GELU -> linear2."""
    B, C, D, H, W = x.shape
    P = patch
    t = x.reshape(B, C, D // P, P, H // P, P, W // P, P).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(B, -1, P * P * P * C)
    x = F.linear(t, sd["vit.patch_embedding.patch_embeddings.1.weight"], sd["vit.patch_embedding.patch_embeddings.1.bias"])
    x = x + sd["vit.patch_embedding.position_embeddings"]
    hidden = []
    hid = x.shape[-1]
    hd = hid // heads
    for i in range(num_layers):
        p = f"vit.blocks.{i}."
        n1 = F.layer_norm(x, (hid,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        qkv = F.linear(n1, sd[p + "attn.qkv.weight"])
        q, k, v = qkv.reshape(B, -1, 3, heads, hd).permute(2, 0, 3, 1, 4)
        att = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1)
        o = (att @ v).permute(0, 2, 1, 3).reshape(B, -1, hid)
        x = x + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        n2 = F.layer_norm(x, (hid,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        m = F.linear(F.gelu(F.linear(n2, sd[p + "mlp.linear1.weight"], sd[p + "mlp.linear1.bias"])),
                     sd[p + "mlp.linear2.weight"], sd[p + "mlp.linear2.bias"])
        x = x + m
        hidden.append(x)
    x = F.layer_norm(x, (hid,), sd["vit.norm.weight"], sd["vit.norm.bias"], 1e-5)
    return x, hidden


def res_block(sd, prefix, x):
    """monai UnetResBlock (kernel 3, stride 1, InstanceNorm3d without affine, LeakyReLU 0.01, convs without bias):
    out = lrelu(norm2(conv2(lrelu(norm1(conv1(x))))) + residual), residual = norm3(conv3(x)) when cin != cout."""
    out = F.conv3d(x, sd[prefix + "conv1.conv.weight"], None, 1, 1)
    out = F.leaky_relu(F.instance_norm(out, eps=1e-5), LRELU)
    out = F.instance_norm(F.conv3d(out, sd[prefix + "conv2.conv.weight"], None, 1, 1), eps=1e-5)
    res = x
    if prefix + "conv3.conv.weight" in sd:
        res = F.instance_norm(F.conv3d(x, sd[prefix + "conv3.conv.weight"], None, 1, 0), eps=1e-5)
    return F.leaky_relu(out + res, LRELU)


def pr_up_block(sd, prefix, x, num_layer):
    """monai UnetrPrUpBlock(conv_block=True, res_block=True): transp_conv_init, then num_layer x [k2s2 transposed conv,
    UnetResBlock]."""
    x = F.conv_transpose3d(x, sd[prefix + "transp_conv_init.conv.weight"], None, 2)
    for j in range(num_layer):
        x = F.conv_transpose3d(x, sd[prefix + f"blocks.{j}.0.conv.weight"], None, 2)
        x = res_block(sd, prefix + f"blocks.{j}.1.", x)
    return x


def up_block(sd, prefix, x, skip):
    """monai UnetrUpBlock: out = transp_conv(x); out = cat((out, skip), 1); out = UnetResBlock(out)."""
    out = F.conv_transpose3d(x, sd[prefix + "transp_conv.conv.weight"], None, 2)
    return res_block(sd, prefix + "conv_block.", torch.cat((out, skip), 1))


def unetr_forward(sd, x_in, heads, num_layers=12):
    """code/networks/unetr.py:215-230 (proj_feat :184-187)."""
    B, _, D, H, W = x_in.shape
    feat = (D // 16, H // 16, W // 16)
    x, hidden = vit_forward(sd, x_in, heads, num_layers)
    hid = x.shape[-1]
    proj = lambda t: t.view(B, feat[0], feat[1], feat[2], hid).permute(0, 4, 1, 2, 3).contiguous()
    enc1 = res_block(sd, "encoder1.layer.", x_in)
    enc2 = pr_up_block(sd, "encoder2.", proj(hidden[3]), 2)
    enc3 = pr_up_block(sd, "encoder3.", proj(hidden[6]), 1)
    enc4 = pr_up_block(sd, "encoder4.", proj(hidden[9]), 0)
    dec3 = up_block(sd, "decoder5.", proj(x), enc4)
    dec2 = up_block(sd, "decoder4.", dec3, enc3)
    dec1 = up_block(sd, "decoder3.", dec2, enc2)
    out = up_block(sd, "decoder2.", dec1, enc1)
    return F.conv3d(out, sd["out.conv.conv.weight"], sd["out.conv.conv.bias"])


def fully_supervised_loss(sd, x, y, heads, n_classes=2, num_layers=12):
    """code/train_fully_supervised_3D_ViT.py: loss = 0.5 * (CE(outputs, label) + Dice(softmax(outputs), label))."""
    from oracle import ssl_oracle as O
    logits = unetr_forward(sd, x, heads, num_layers)
    loss, ce, dice = O.supervised_loss(logits, y, n_classes)
    return loss, logits
