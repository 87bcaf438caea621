"""Golden vectors for the activation extractor from the LIVE reference (build container only).
Writes tests/golden/hooks.pt.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_hooks.py

A tiny random-init Whisper (seeded; no checkpoint is available offline) is pushed through the
reference's ``extract_features_batch`` on the CPU.  Stored: the recipe to rebuild the identical model
and input, the RAW hooked hidden states (from ``output_hidden_states``), and the reference's
extracted encoder activations.  The reference's decoder hook is broken under the installed
transformers (it indexes ``output[0]`` of a bare tensor, hooks.py:101), so the decoder expectation
is its documented intent: the decoder's final LayerNorm applied to the layer output.
"""

from __future__ import annotations

import sys
from pathlib import Path

import torch

REF_SRC = Path("/root/reference/src")
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF_SRC))
sys.dont_write_bytecode = True

from transformers import WhisperConfig, WhisperForConditionalGeneration  # noqa: E402

from whisper_sae.sae.hooks import extract_features_batch, flatten_activations  # noqa: E402  (reference)

GOLDEN = ROOT / "tests" / "golden"
RECIPE = dict(d_model=64, encoder_layers=3, decoder_layers=2, encoder_attention_heads=2,
              decoder_attention_heads=2, encoder_ffn_dim=128, decoder_ffn_dim=128, num_mel_bins=80,
              max_source_positions=40, max_target_positions=16, vocab_size=200, decoder_start_token_id=1,
              pad_token_id=0, bos_token_id=1, eos_token_id=2, model_seed=11, input_seed=12, batch=3)


def build(recipe: dict):
    cfg = WhisperConfig(**{k: v for k, v in recipe.items() if k not in ("model_seed", "input_seed", "batch")})
    torch.manual_seed(recipe["model_seed"])
    model = WhisperForConditionalGeneration(cfg).eval()
    with torch.no_grad():     # give the final LayerNorms non-trivial affine parameters
        g = torch.Generator().manual_seed(recipe["model_seed"] + 1)
        for ln in (model.model.encoder.layer_norm, model.model.decoder.layer_norm):
            ln.weight.copy_(1.0 + 0.1 * torch.randn(ln.weight.shape, generator=g))
            ln.bias.copy_(0.1 * torch.randn(ln.bias.shape, generator=g))
    x = torch.randn(recipe["batch"], recipe["num_mel_bins"], 2 * recipe["max_source_positions"],
                    generator=torch.Generator().manual_seed(recipe["input_seed"]))
    return model, x


def main() -> None:
    model, x = build(RECIPE)
    enc_layers = [0, 2]
    ref = extract_features_batch(model, x, encoder_layers=enc_layers, decoder_layers=[], device="cpu")
    with torch.no_grad():
        enc_out = model.model.encoder(x, output_hidden_states=True)
        start = torch.full((x.size(0), 1), model.config.decoder_start_token_id, dtype=torch.long)
        dec_out = model.model.decoder(input_ids=start, encoder_hidden_states=enc_out.last_hidden_state,
                                      output_hidden_states=True)
        # hidden_states[l + 1] is the output of layer l ... except that the LAST entry already went
        # through the final LayerNorm; take the raw layer outputs with hooks instead
        raw_enc, raw_dec = {}, {}
        hs = [model.model.encoder.layers[l].register_forward_hook(
            lambda m, i, o, l=l: raw_enc.__setitem__(l, (o[0] if isinstance(o, tuple) else o).clone()))
            for l in range(RECIPE["encoder_layers"])]
        hs += [model.model.decoder.layers[l].register_forward_hook(
            lambda m, i, o, l=l: raw_dec.__setitem__(l, (o[0] if isinstance(o, tuple) else o).clone()))
            for l in range(RECIPE["decoder_layers"])]
        e = model.model.encoder(x).last_hidden_state
        model.model.decoder(input_ids=start, encoder_hidden_states=e)
        for h in hs:
            h.remove()
        dec_expected = {l: model.model.decoder.layer_norm(raw_dec[l]) for l in raw_dec}
    fixture = {
        "torch": torch.__version__, "recipe": RECIPE, "encoder_layers": enc_layers,
        "raw_encoder": raw_enc, "raw_decoder": raw_dec,
        "encoder_ln": (model.model.encoder.layer_norm.weight.detach().clone(),
                       model.model.encoder.layer_norm.bias.detach().clone(), model.model.encoder.layer_norm.eps),
        "decoder_ln": (model.model.decoder.layer_norm.weight.detach().clone(),
                       model.model.decoder.layer_norm.bias.detach().clone(), model.model.decoder.layer_norm.eps),
        "reference_encoder": {l: ref["encoder"][l].clone() for l in enc_layers},
        "reference_encoder_flat": {l: flatten_activations(ref["encoder"][l], "encoder").clone() for l in enc_layers},
        "decoder_expected": dec_expected,
    }
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.save(fixture, GOLDEN / "hooks.pt")
    for l in enc_layers:
        print("encoder layer", l, tuple(ref["encoder"][l].shape), float(ref["encoder"][l].abs().mean()))
    print("decoder layers", {l: tuple(v.shape) for l, v in dec_expected.items()})


if __name__ == "__main__":
    main()
