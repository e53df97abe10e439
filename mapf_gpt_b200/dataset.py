"""Training-side reuse of the forward path (first slice): the reference's Arrow shards evaluated by the B200 engine.

  * `MapfArrowDataset(folder_path, device, batch_size)` -- same constructor and iteration contract as
    dataset/fast_data_loader.py:13-71: `*.arrow` files of schema `input_tensors: list<int8>[256]`, `gt_actions: int8`
    (dataset/generate_dataset.py:188-191), files of a `train` folder split over `LOCAL_RANK` / `WORLD_SIZE` (`:21-28`), rows
    shuffled inside a file (`:45-47`), batches `(input_tensors [B, 256] int8, target_tensors [B, 256] int8)` with -1 everywhere
    except the last position, which carries the ground-truth action (`:57`).  Host numpy arrays (`device` is kept for signature
    compatibility: the engine takes host buffers through the C ABI).
  * `estimate_loss(engine, data_iter, eval_iters)` -- train.py:244-258 for one split: the mean over `eval_iters` batches of the
    batch-mean cross-entropy (model.py:180-183, ignore_index -1), plus the action accuracy the paper reports.
  * `write_shard(path, input_tensors, gt_actions)` -- a shard in the reference's format (tests, and rollouts logged by the engine).

No backward pass here: fine-tuning from rollouts would need one (not built)."""
from __future__ import annotations

import glob
import os

import numpy as np


def _pa():
    import pyarrow as pa
    import pyarrow.ipc  # noqa: F401
    return pa


def write_shard(path, input_tensors, gt_actions) -> None:
    """generate_dataset.py:188-211: one Arrow IPC file, `input_tensors` list<int8>, `gt_actions` int8."""
    pa = _pa()
    x = np.ascontiguousarray(input_tensors, dtype=np.int8)
    y = np.ascontiguousarray(gt_actions, dtype=np.int8)
    assert x.ndim == 2 and y.shape == (x.shape[0],)
    schema = pa.schema([("input_tensors", pa.list_(pa.int8())), ("gt_actions", pa.int8())])
    offsets = pa.array(np.arange(0, x.size + 1, x.shape[1], dtype=np.int32))
    col = pa.ListArray.from_arrays(offsets, pa.array(x.reshape(-1)))
    table = pa.Table.from_arrays([col, pa.array(y)], schema=schema)
    with open(path, "wb") as f:
        with pa.ipc.new_file(f, schema) as writer:
            writer.write(table)


def read_shard(path):
    """-> (input_tensors int8 [n, T], gt_actions int8 [n]) of one shard (fast_data_loader.py:38-43)."""
    pa = _pa()
    with pa.memory_map(str(path)) as source:
        table = pa.ipc.open_file(source).read_all()
    col = table["input_tensors"].combine_chunks()
    flat = col.flatten().to_numpy(zero_copy_only=False)
    n = len(col)
    if n == 0:
        return np.zeros((0, 256), np.int8), np.zeros((0,), np.int8)
    x = np.ascontiguousarray(flat, dtype=np.int8).reshape(n, -1)
    y = np.ascontiguousarray(table["gt_actions"].to_numpy(zero_copy_only=False), dtype=np.int8)
    return x, y


class MapfArrowDataset:
    def __init__(self, folder_path, device=None, batch_size: int = 2048, seed=None):
        self.all_data_files = self.file_paths = sorted(glob.glob(os.path.join(str(folder_path), "*.arrow")))
        if not self.file_paths:
            raise FileNotFoundError(f"no *.arrow shards under {folder_path}")
        self.device, self.batch_size = device, batch_size
        self._rng = np.random.default_rng(seed)
        rank, world = os.environ.get("LOCAL_RANK"), os.environ.get("WORLD_SIZE")
        if "train" in str(folder_path) and rank is not None and world is not None:   # fast_data_loader.py:21-28
            rank, world = int(rank), int(world)
            per = len(self.file_paths) // world
            self.file_paths = self.file_paths[rank * per:(rank + 1) * per]
        self._rows_per_file = len(read_shard(self.file_paths[0])[1])

    def _load(self, path):
        x, y = read_shard(path)
        idx = self._rng.permutation(len(x))                      # shuffle inside the file (fast_data_loader.py:45-47)
        x, y = x[idx], y[idx]
        t = np.full(x.shape, -1, dtype=np.int8)
        t[:, -1] = y
        return x, t

    def __iter__(self):
        while True:
            for path in self.file_paths:
                x, t = self._load(path)
                for i in range(0, len(x), self.batch_size):
                    yield x[i:i + self.batch_size], t[i:i + self.batch_size]

    def get_shard_size(self):
        return self._rows_per_file * len(self.file_paths)

    def get_full_dataset_size(self):
        return self._rows_per_file * len(self.all_data_files)


def estimate_loss(engine, data_iter, eval_iters: int = 40) -> dict:
    """train.py:244-258 for one split, on the engine: {"loss": mean of the batch-mean cross-entropies, "accuracy": share of rows
    whose arg-max action equals the ground truth, "rows": rows evaluated}."""
    losses, hits, rows = [], 0, 0
    for _ in range(eval_iters):
        x, t = next(data_iter)
        y = t[:, -1]
        loss, pred = engine.eval_tokens(x, y)
        valid = y >= 0
        losses.append(float(loss[valid].mean()) if valid.any() else 0.0)     # F.cross_entropy: mean over non-ignored targets
        hits += int((pred[valid] == y[valid]).sum())
        rows += int(valid.sum())
    return {"loss": float(np.mean(losses)), "accuracy": hits / max(rows, 1), "rows": rows}


def main():
    """python -m mapf_gpt_b200.dataset --folder dataset/validation --weights weights/MAPF-GPT-2M.pt [--eval_iters 40]"""
    import argparse
    import json
    from pathlib import Path

    from . import engine as E, weights as W
    ap = argparse.ArgumentParser(description="validation loss / action accuracy of a checkpoint on Arrow shards (train.py:244-258)")
    ap.add_argument("--folder", required=True)
    ap.add_argument("--weights", default="weights/MAPF-GPT-2M.pt")
    ap.add_argument("--model", default=None, choices=[None, "2M", "6M", "85M"], help="seeded random init of this size when --weights is missing")
    ap.add_argument("--batch_size", type=int, default=4096)
    ap.add_argument("--eval_iters", type=int, default=40)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    if Path(a.weights).exists():
        sd, cfg = W.load_checkpoint(a.weights)
        src = a.weights
    elif a.model:
        cfg = W.model_config(a.model)
        sd = W.random_init(cfg, 1234)
        src = f"seeded random init ({a.model})"
    else:
        raise SystemExit(f"{a.weights} not found (pass --model for a random init)")
    eng = E.RolloutEngine(1, 1, 11, 11, device=a.device)
    eng.load_model(sd, cfg)
    ds = MapfArrowDataset(a.folder, device=f"cuda:{a.device}", batch_size=a.batch_size)
    out = estimate_loss(eng, iter(ds), a.eval_iters)
    out.update(weights=src, folder=a.folder, files=len(ds.file_paths))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
