#!/usr/bin/env python
"""Mint golden fixtures from the REAL reference model class.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

Imports ``TransducerGRU`` / ``ModelHandler`` unmodified from the reference checkout
(``helen/modules/python/models/TransducerModel.py``, ``ModelHander.py``) and drives them
with a literal transcription of the reference driver loop
(``helen/modules/python/models/predict.py:90-154``) -- kept here, outside the product and
outside ``oracle/``, so that the fixtures do not depend on the code they pin.

Outputs (committed): ``tests/golden/model_*.npz`` (state_dicts) and
``tests/golden/case_*.npz`` (inputs + reference outputs in fp32 and fp64).
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REFERENCE = os.environ.get("HELEN_REFERENCE", "/root/reference")
sys.path.insert(0, REFERENCE)
from helen.modules.python.models.TransducerModel import TransducerGRU  # noqa: E402
from helen.modules.python.models.ModelHander import ModelHandler  # noqa: E402
from helen.modules.python.Options import ImageSizeOptions, TrainOptions  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_predict(model, images_u8, seq_length):
    """predict.py:90-154 with SEQ_LENGTH replaced by the image's own length (the
    reference hard-codes 1000 == images.size(1))."""
    dtype = next(model.parameters()).dtype
    with torch.no_grad():
        images = images_u8.type(torch.FloatTensor).to(dtype)
        hidden = torch.zeros(images.size(0), 2 * TrainOptions.GRU_LAYERS, TrainOptions.HIDDEN_SIZE, dtype=dtype)
        prediction_base_tensor = torch.zeros((images.size(0), images.size(1), ImageSizeOptions.TOTAL_BASE_LABELS), dtype=dtype)
        prediction_rle_tensor = torch.zeros((images.size(0), images.size(1), ImageSizeOptions.TOTAL_RLE_LABELS), dtype=dtype)
        first_chunk = None
        for i in range(0, seq_length, TrainOptions.WINDOW_JUMP):
            if i + TrainOptions.TRAIN_WINDOW > seq_length:
                break
            chunk_start = i
            chunk_end = i + TrainOptions.TRAIN_WINDOW
            image_chunk = images[:, chunk_start:chunk_end]
            output_base, output_rle, hidden = model(image_chunk, hidden)
            if first_chunk is None:
                first_chunk = (output_base.clone(), output_rle.clone(), hidden.clone())
            top_zeros = chunk_start
            bottom_zeros = seq_length - chunk_end
            inference_layers = nn.Sequential(nn.Softmax(dim=2), nn.ZeroPad2d((0, 0, top_zeros, bottom_zeros)))
            base_prediction = inference_layers(output_base)
            rle_prediction = inference_layers(output_rle)
            prediction_base_tensor = torch.add(prediction_base_tensor, base_prediction)
            prediction_rle_tensor = torch.add(prediction_rle_tensor, rle_prediction)
        base_values, base_labels = torch.max(prediction_base_tensor, 2)
        rle_values, rle_labels = torch.max(prediction_rle_tensor, 2)
    return {
        "base_prob": prediction_base_tensor.numpy(),
        "rle_prob": prediction_rle_tensor.numpy(),
        "base_label": base_labels.numpy().astype(np.uint8),
        "rle_label": rle_labels.numpy().astype(np.uint8),
        "hidden": hidden.numpy(),
        "chunk0_base": first_chunk[0].numpy(),
        "chunk0_rle": first_chunk[1].numpy(),
        "chunk0_hidden": first_chunk[2].numpy(),
    }


def new_model(features, seed, scale=1.0):
    torch.manual_seed(seed)
    model = ModelHandler.get_new_gru_model(input_channels=1, image_features=features, gru_layers=1,
                                           hidden_size=128, num_base_classes=5, num_rle_classes=11)
    if scale != 1.0:
        with torch.no_grad():
            for p in model.parameters():
                p.mul_(scale)
    return model.eval()


def uniform_images(batch, seq, features, seed):
    gen = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, seq, features), dtype=torch.uint8, generator=gen)


def pileup_images(batch, seq, features, seed):
    """Sparse pileup-like columns: 1-3 active features per column, weights summing to 255."""
    rng = np.random.default_rng(seed)
    img = np.zeros((batch, seq, features), dtype=np.uint8)
    for b in range(batch):
        for t in range(seq):
            k = int(rng.integers(1, 4))
            idx = rng.choice(features, size=k, replace=False)
            w = rng.dirichlet(np.ones(k))
            img[b, t, idx] = np.round(255 * w).astype(np.uint8)
    return torch.from_numpy(img)


def ragged_images(batch, seq, features, seed):
    """Right-padded short images (dataloader_predict.py:74-82 pads with zero columns)."""
    img = uniform_images(batch, seq, features, seed).clone()
    lengths = [seq, max(1, seq // 3), 0, seq - 1][:batch]
    for b, n in enumerate(lengths):
        img[b, n:] = 0
    return img


MODELS = {
    # name: (features, seed, scale)
    "F10_seed0": (10, 0, 1.0),
    "F90_seed0": (90, 0, 1.0),
    "F90_sharp": (90, 2, 2.5),   # wider logit margins, saturating gates ("trained-like")
}

CASES = {
    # name: (model, image maker, batch, seq, image seed)
    "cfg1_F10_B64_T100": ("F10_seed0", uniform_images, 64, 100, 1),      # BASELINE config 1 (64 windows)
    "F10_B4_T1000_uniform": ("F10_seed0", uniform_images, 4, 1000, 1),
    "F10_B3_T1000_pileup": ("F10_seed0", pileup_images, 3, 1000, 3),
    "F90_B3_T1000_uniform": ("F90_seed0", uniform_images, 3, 1000, 1),
    "F90_B4_T1000_ragged": ("F90_seed0", ragged_images, 4, 1000, 5),
    "F90_B2_T150_uniform": ("F90_seed0", uniform_images, 2, 150, 7),     # 2 chunks, overlap 50..100
    "F90_B1_T100_pileup": ("F90_seed0", pileup_images, 1, 100, 9),
    "F90sharp_B3_T1000_pileup": ("F90_sharp", pileup_images, 3, 1000, 11),
    "F90sharp_B2_T1000_uniform": ("F90_sharp", uniform_images, 2, 1000, 13),
}


def main():
    torch.set_num_threads(1)
    models = {}
    for name, (features, seed, scale) in MODELS.items():
        model = new_model(features, seed, scale)
        models[name] = model
        sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
        np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), **sd)
        print("model", name, sum(v.size for v in sd.values()), "params")
    for name, (model_name, maker, batch, seq, seed) in CASES.items():
        model = models[model_name]
        images = maker(batch, seq, model.gru_encoder.input_size, seed)
        out32 = reference_predict(model, images, seq)
        model64 = new_model(*MODELS[model_name]).double()
        out64 = reference_predict(model64, images, seq)
        payload = {"images": images.numpy(), "model": np.array(model_name),
                   "torch_version": np.array(torch.__version__)}
        payload.update({f"f32_{k}": v for k, v in out32.items()})
        for k in ("base_prob", "rle_prob", "base_label", "rle_label"):
            payload[f"f64_{k}"] = out64[k]
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **payload)
        flips = int((out32["base_label"] != out64["base_label"]).sum() + (out32["rle_label"] != out64["rle_label"]).sum())
        print("case", name, "fp32-vs-fp64 label flips:", flips,
              "max|dP|", float(np.abs(out32["base_prob"] - out64["base_prob"]).max()))


if __name__ == "__main__":
    main()
