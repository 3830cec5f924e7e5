"""Minimal host equivalents of bm25s/utils/corpus.py used by save/load: newline index + lazy JSONL reader."""
import json
import os

import numpy as np


def change_extension(path, new_extension):
    path = str(path)
    return path.rpartition(".")[0] + new_extension


def find_newline_positions(path, show_progress=True, leave_progress=True):
    """Byte offset of the start of every line (bm25s/utils/corpus.py, find_newline_positions)."""
    indexes = []
    with open(path, "r", encoding="utf-8") as f:
        indexes.append(f.tell())
        while f.readline():
            indexes.append(f.tell())
    return indexes[:-1]


def save_mmindex(indexes, path):
    with open(change_extension(path, ".mmindex.json"), "w") as f:
        f.write(json.dumps(indexes))


def load_mmindex(path):
    with open(change_extension(path, ".mmindex.json"), "r") as f:
        return json.loads(f.read())


class JsonlCorpus:
    """Lazy, index-addressable view of a corpus.jsonl file (bm25s/utils/corpus.py:105-220)."""

    def __init__(self, path, show_progress=True, leave_progress=True, save_index=True, verbosity=1):
        self.path = str(path)
        mm = change_extension(self.path, ".mmindex.json")
        if os.path.exists(mm):
            self.mmindex = load_mmindex(self.path)
        else:
            self.mmindex = find_newline_positions(self.path)
            if save_index:
                save_mmindex(self.mmindex, self.path)
        self.file_obj = open(self.path, "r", encoding="utf-8")

    def __len__(self):
        return len(self.mmindex)

    def _get(self, i):
        self.file_obj.seek(self.mmindex[int(i)])
        return json.loads(self.file_obj.readline())

    def __getitem__(self, index):
        if isinstance(index, (int, np.integer)):
            return self._get(index)
        if isinstance(index, slice):
            return [self._get(i) for i in range(*index.indices(len(self)))]
        if isinstance(index, (list, tuple)):
            return [self._get(i) for i in index]
        if isinstance(index, np.ndarray):
            flat = [self._get(i) for i in index.flatten().tolist()]
            return np.array(flat, dtype=object).reshape(index.shape)
        raise TypeError("Invalid index type")

    def close(self):
        if getattr(self, "file_obj", None) is not None:
            self.file_obj.close()
            self.file_obj = None

    def __del__(self):
        self.close()
