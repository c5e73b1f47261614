"""GPU: the example entry points that stand where src/run_mim.py and src/run_inference.py / run_inspect.py stand
(small model, synthetic int16 volumes): the loop trains (loss falls), checkpoints load into the upstream class, the
extractor writes reference-format parquet / npy files, resumes, and matches a direct model call."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

import __graft_entry__ as ge

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
OVR = ",".join(f"{k}={ge.SMALL64[k]}" for k in ["hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "decoder_hidden_size",
                                              "decoder_num_hidden_layers", "decoder_num_attention_heads", "decoder_intermediate_size"])


def test_train_mim_example_learns_and_checkpoints(tmp_path):
    import transformers

    import train_mim

    out = str(tmp_path / "ckpt")
    losses = train_mim.main(["--synthetic", "2", "--image_size", "96", "--depth", "96", "--steps", "12", "--batch", "2", "--learning_rate", "1e-3",
                             "--warmup_ratio", "0.1", "--config_overrides", OVR, "--output_dir", out])
    assert len(losses) == 12 and all(np.isfinite(losses)) and losses[-1] < losses[0]
    up = transformers.VideoMAEForPreTraining.from_pretrained(out)  # the reference's class loads what we saved
    assert up.config.hidden_size == 128 and os.path.exists(os.path.join(out, "optimizer.pt"))


def test_extract_embeddings_example_formats_and_resume(tmp_path):
    import extract_embeddings
    from smb_vision_b200.data import VolumePreprocessor
    from smb_vision_b200.modeling import B200VideoMAEModel

    save = str(tmp_path / "emb")
    common = ["--image_size", "96", "--depth", "96", "--config_overrides", OVR, "--save_dir", save]
    files = extract_embeddings.main(["--synthetic", "5", "--format", "parquet", "--model_id", "small64"] + common)
    assert len(files) == 5
    df = pd.read_parquet(os.path.join(save, "model_id=small64", "synthetic_0003.parquet"))
    assert list(df["embedding_shape"][0]) == [216, 128] and df["uid"][0] == "synthetic_0003"
    # same numbers as a direct call on the same prepared volume
    g = torch.Generator().manual_seed(0)
    raws = [torch.randint(-1100, 1500, (96, 96, 96), generator=g, dtype=torch.int16) for _ in range(5)]
    hc = ge.hf_config(ge.SMALL64)
    torch.manual_seed(0)
    model = B200VideoMAEModel(hc).to("cuda:0").eval()
    x = VolumePreprocessor(96, 96, device="cuda:0")(raws[3]).unsqueeze(0)
    want = model(x).last_hidden_state[0].cpu().numpy()
    assert np.array_equal(np.asarray(df["embedding"][0]).reshape(216, 128), want)
    assert extract_embeddings.main(["--synthetic", "5", "--format", "parquet", "--model_id", "small64"] + common) == []  # resume: nothing left
    npys = extract_embeddings.main(["--synthetic", "2", "--format", "npy", "--save_dir", str(tmp_path / "npy")] + common[:-2])
    assert [os.path.basename(p) for p in npys] == ["synthetic_0000.npy", "synthetic_0001.npy"]
    assert np.load(npys[0]).shape == (1, 216, 128)
