"""Input pipeline without LMDB / pyarrow-0.14.

The reference stores pre-cut clips in an LMDB of ``pyarrow.serialize``d tuples (an API that no longer exists in
current pyarrow) and turns one into a training sample in ``SpeechMotionDataset.__getitem__``
(scripts/data_loader/lmdb_data_loader_expressive.py:108-176; TED-Gesture twin lmdb_data_loader.py).  This module keeps
the per-sample logic bit-exact -- word-to-frame placement, SOS/EOS word tensor, clipping of audio / spectrogram / pose
to ``n_poses`` frames -- over a flat shard format (one ``.npz`` of padded arrays + offsets per shard), and returns the
same 7-tuple, so ``default_collate_fn`` and the training loop (train_expressive.py:321-337) work unchanged.

Index contract (bit-exact): ``extend_word_seq`` puts the id of each word at frame
``max(0, floor((word_start - clip_start) / frame_duration))`` when that frame is < n_poses (later words overwrite earlier
ones on the same frame); with ``remove_word_timing`` the words are spread evenly at ``(i+1) * int(n_poses / (n+1))``.
"""
from __future__ import annotations

import json
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset
from torch.utils.data.dataloader import default_collate


def calc_spectrogram_length_from_motion_length(n_frames: int, fps: float) -> int:
    """utils/data_utils_expressive.py:91-93."""
    return int(round((n_frames / fps * 16000 - 1024) / 512 + 1))


def make_audio_fixed_length(audio: np.ndarray, expected_audio_length: int) -> np.ndarray:
    """utils/data_utils_expressive.py:118-124 (symmetric padding or truncation)."""
    n_padding = expected_audio_length - len(audio)
    if n_padding > 0:
        return np.pad(audio, (0, n_padding), mode="symmetric")
    return audio[0:expected_audio_length]


def extend_word_seq(lang, words: Sequence, start_time: float, end_time: float, n_frames: int,
                    remove_word_timing: bool = False) -> torch.Tensor:
    """lmdb_data_loader_expressive.py:116-141: per-frame word ids (0 = PAD) of one clip."""
    frame_duration = (end_time - start_time) / n_frames
    ext = np.zeros(n_frames)
    if remove_word_timing:
        n_words = 0
        for word in words:
            idx = max(0, int(np.floor((word[1] - start_time) / frame_duration)))
            if idx < n_frames:
                n_words += 1
        space = int(n_frames / (n_words + 1))
        for i in range(n_words):
            ext[(i + 1) * space] = lang.get_word_index(words[i][0])
    else:
        for word in words:
            idx = max(0, int(np.floor((word[1] - start_time) / frame_duration)))
            if idx < n_frames:
                ext[idx] = lang.get_word_index(word[0])
    return torch.Tensor(ext).long()


def words_to_tensor(lang, words: Sequence, end_time=None) -> torch.Tensor:
    """lmdb_data_loader_expressive.py:143-150: [SOS, ids of the words that start before end_time..., EOS]."""
    indexes = [lang.SOS_token]
    for word in words:
        if end_time is not None and word[1] > end_time:
            break
        indexes.append(lang.get_word_index(word[0]))
    indexes.append(lang.EOS_token)
    return torch.Tensor(indexes).long()


def default_collate_fn(data):
    """lmdb_data_loader_expressive.py:45-55."""
    _, text_padded, pose_seq, vec_seq, audio, spectrogram, aux_info = zip(*data)
    aux = {key: default_collate([d[key] for d in aux_info]) for key in aux_info[0]}
    return (torch.tensor([0]), torch.tensor([0]), default_collate(text_padded), default_collate(pose_seq),
            default_collate(vec_seq), default_collate(audio), default_collate(spectrogram), aux)


def write_shard(path: str, samples: List[Tuple]) -> None:
    """samples: list of (word_seq [(word, start, end)...], pose_seq [F,J,3], vec_seq [F,J-1,3] or [F,D], audio [n],
    spectrogram [128,S], aux_info {'vid', 'start_time', 'end_time', ...}) -- the tuple the reference's DataPreprocessor
    stores per clip (data_preprocessor_expressive.py:144-160).  Arrays are concatenated with offset tables."""
    cat = lambda xs: np.concatenate([np.asarray(x).reshape(-1) for x in xs]) if xs else np.zeros(0)
    off = lambda xs: np.cumsum([0] + [int(np.asarray(x).size) for x in xs]).astype(np.int64)
    pose = [np.asarray(s[1], dtype=np.float32) for s in samples]
    vec = [np.asarray(s[2], dtype=np.float32) for s in samples]
    audio = [np.asarray(s[3], dtype=np.float32) for s in samples]
    spec = [np.asarray(s[4], dtype=np.float16) for s in samples]
    meta = [{"words": [[w[0], float(w[1]), float(w[2])] for w in s[0]], "aux": s[5],
             "pose_shape": list(pose[i].shape), "vec_shape": list(vec[i].shape), "spec_shape": list(spec[i].shape)}
            for i, s in enumerate(samples)]
    np.savez(path, pose=cat(pose).astype(np.float32), pose_off=off(pose), vec=cat(vec).astype(np.float32), vec_off=off(vec),
             audio=cat(audio).astype(np.float32), audio_off=off(audio), spec=cat(spec).astype(np.float16), spec_off=off(spec),
             meta=np.frombuffer(json.dumps(meta, default=_json_default).encode(), dtype=np.uint8))


def _json_default(o):
    """aux_info values arrive as numpy scalars / arrays from the preprocessor's arithmetic."""
    if isinstance(o, np.generic):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    raise TypeError(f"Object of type {type(o).__name__} is not JSON serializable")


class SpeechMotionShardDataset(Dataset):
    """Same ``__getitem__`` contract as the reference's ``SpeechMotionDataset`` over one flat shard."""

    def __init__(self, shard_path: str, n_poses: int, pose_resampling_fps: float, speaker_model=None,
                 remove_word_timing: bool = False):
        z = np.load(shard_path if shard_path.endswith(".npz") else shard_path + ".npz")
        self._z = {k: z[k] for k in z.files}
        self._meta = json.loads(bytes(self._z["meta"]).decode())
        self.n_samples = len(self._meta)
        self.n_poses = n_poses
        self.remove_word_timing = remove_word_timing
        self.expected_audio_length = int(round(n_poses / pose_resampling_fps * 16000))
        self.expected_spectrogram_length = calc_spectrogram_length_from_motion_length(n_poses, pose_resampling_fps)
        self.lang_model = None
        self.speaker_model = speaker_model

    def set_lang_model(self, lang_model):
        self.lang_model = lang_model

    def __len__(self):
        return self.n_samples

    def _arr(self, name: str, idx: int, shape) -> np.ndarray:
        o = self._z[name + "_off"]
        return self._z[name][o[idx]:o[idx + 1]].reshape(shape)

    def __getitem__(self, idx: int):
        m = self._meta[idx]
        word_seq, aux_info = m["words"], dict(m["aux"])
        pose_seq = self._arr("pose", idx, m["pose_shape"])
        vec_seq = self._arr("vec", idx, m["vec_shape"])
        audio = self._arr("audio", idx, (-1,))
        spectrogram = self._arr("spec", idx, m["spec_shape"])
        duration = aux_info["end_time"] - aux_info["start_time"]
        # clipping to n_poses frames (lmdb_data_loader_expressive.py:152-160)
        sample_end_time = aux_info["start_time"] + duration * self.n_poses / vec_seq.shape[0]
        audio = make_audio_fixed_length(audio, self.expected_audio_length)
        spectrogram = spectrogram[:, 0:self.expected_spectrogram_length]
        vec_seq = vec_seq[0:self.n_poses]
        pose_seq = pose_seq[0:self.n_poses]
        word_seq_tensor = words_to_tensor(self.lang_model, word_seq, sample_end_time)
        extended = extend_word_seq(self.lang_model, word_seq, aux_info["start_time"], sample_end_time, self.n_poses,
                                   self.remove_word_timing)
        vec_t = torch.from_numpy(np.array(vec_seq)).reshape((vec_seq.shape[0], -1)).float()
        pose_t = torch.from_numpy(np.array(pose_seq)).reshape((pose_seq.shape[0], -1)).float()
        return (word_seq_tensor, extended, pose_t, vec_t, torch.from_numpy(np.array(audio)).float(),
                torch.from_numpy(np.array(spectrogram)).float(), aux_info)
