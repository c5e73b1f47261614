"""CPU: embedding output formats against the reference's own writer calls, re-run here
(np.save / pandas.DataFrame.to_parquet exactly as src/run_inference.py:89-96 and run_inspect.py:140-175 do)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from smb_vision_b200.output import EmbeddingWriter, npy_stem, processed_uids, unprocessed_files, write_parquet


def _reference_parquet(path, emb_last_hidden_state, uid, model_id):
    # run_inspect.py:150-172, verbatim semantics
    np_embedding = emb_last_hidden_state.squeeze(0).float().cpu().numpy()
    df = pd.DataFrame({"uid": [uid], "embedding": [np_embedding.flatten()], "embedding_shape": [np_embedding.shape], "model_id": [model_id]})
    df.to_parquet(path, compression="snappy")


def test_parquet_equals_reference_writer(tmp_path):
    emb = torch.randn(1, 216, 128, generator=torch.Generator().manual_seed(0))
    ref = tmp_path / "ref.parquet"
    _reference_parquet(ref, emb, "uid_7", "smb-vision-base")
    out = write_parquet(str(tmp_path / "out"), "uid_7", emb.numpy(), "smb-vision-base")
    assert out.endswith(os.path.join("model_id=smb-vision-base", "uid_7.parquet"))
    a, b = pd.read_parquet(ref), pd.read_parquet(out)
    assert list(a.columns) == list(b.columns) == ["uid", "embedding", "embedding_shape", "model_id"]
    assert a["uid"][0] == b["uid"][0] and a["model_id"][0] == b["model_id"][0]
    assert np.array_equal(np.asarray(a["embedding"][0]), np.asarray(b["embedding"][0])) and np.asarray(b["embedding"][0]).dtype == np.float32
    assert list(a["embedding_shape"][0]) == list(b["embedding_shape"][0]) == [216, 128]
    back = np.asarray(b["embedding"][0]).reshape(list(b["embedding_shape"][0]))
    assert np.array_equal(back, emb[0].numpy())


def test_npy_equals_reference_writer(tmp_path):
    emb = torch.randn(1, 72, 64, generator=torch.Generator().manual_seed(1))
    w = EmbeddingWriter(str(tmp_path), fmt="npy")
    w.submit("/data/ct_0001.nii.gz", emb)
    (p,) = w.close()
    assert os.path.basename(p) == "ct_0001.npy" and npy_stem("a/b/scan.nii") == "scan"
    ref = tmp_path / "ref.npy"
    np.save(ref, emb.cpu().numpy())  # run_inference.py:91-92
    assert open(p, "rb").read() == open(ref, "rb").read()  # byte-identical file


def test_async_writer_and_resume(tmp_path):
    save = str(tmp_path / "emb")
    img = tmp_path / "img"
    img.mkdir()
    for i in range(5):
        (img / f"u{i}.nii.gz").write_bytes(b"")
    (img / "notes.txt").write_bytes(b"")
    assert sorted(f["uid"] for f in unprocessed_files(str(img), save)) == [f"u{i}" for i in range(5)]
    buf = torch.empty(1, 8, 4)  # one reused host buffer, like the runner's pinned ring: the writer must copy
    w = EmbeddingWriter(save, fmt="parquet", model_id="m1", workers=2, max_pending=2)
    for i in range(3):
        buf.fill_(float(i))
        w.submit(f"u{i}", buf)
    files = w.close()
    assert len(files) == 3 and processed_uids(save) == {"u0", "u1", "u2"}
    for i in range(3):
        df = pd.read_parquet(os.path.join(save, "model_id=m1", f"u{i}.parquet"))
        assert float(np.asarray(df["embedding"][0])[0]) == float(i) and list(df["embedding_shape"][0]) == [8, 4]
    assert sorted(f["uid"] for f in unprocessed_files(str(img), save)) == ["u3", "u4"]  # resume skips what exists
    assert not [f for f in os.listdir(os.path.join(save, "model_id=m1")) if f.endswith(".tmp")]
    with pytest.raises(ValueError):
        EmbeddingWriter(save, fmt="parquet")
