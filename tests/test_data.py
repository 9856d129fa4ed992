"""Input pipeline: the per-sample logic of SpeechMotionDataset.__getitem__
(lmdb_data_loader_expressive.py:108-176) over the flat shard format -- index contracts checked against values
worked out by hand from the reference's formulas, shard round trip, collate shapes."""
import numpy as np
import torch

from ha2g_b200 import data
from ha2g_b200.model.vocab import Vocab


def _lang():
    lang = Vocab("words")
    for w in ("hello", "big", "world"):
        lang.index_word(w)
    return lang


def test_spectrogram_and_audio_lengths():
    assert data.calc_spectrogram_length_from_motion_length(34, 15) == 70
    a = np.arange(5, dtype=np.float32)
    assert np.array_equal(data.make_audio_fixed_length(a, 3), a[:3])
    assert np.array_equal(data.make_audio_fixed_length(a, 8), np.array([0, 1, 2, 3, 4, 4, 3, 2], dtype=np.float32))


def test_extend_word_seq_index_contract():
    lang = _lang()
    words = [["hello", 10.05, 10.3], ["unk-word", 10.9, 11.0], ["big", 11.0, 11.2], ["world", 12.2, 12.4], ["late", 13.0, 13.1]]
    start, end, n = 10.0, 10.0 + 34 / 15, 34              # frame_duration = 1/15 s
    ext = data.extend_word_seq(lang, words, start, end, n)
    fd = (end - start) / n
    want = np.zeros(n, dtype=np.int64)
    for w, s, _ in words:
        i = max(0, int(np.floor((s - start) / fd)))
        if i < n:
            want[i] = lang.get_word_index(w)
    assert ext.dtype == torch.int64 and np.array_equal(ext.numpy(), want)
    assert ext[0] == lang.get_word_index("hello") and ext[13] == Vocab.UNK_token and ext[15] == lang.get_word_index("big")
    assert ext[33] == lang.get_word_index("world") and int((ext != 0).sum()) == 4      # "late" falls outside
    # a word that starts before the clip lands on frame 0; a later word on the same frame overwrites it
    ext2 = data.extend_word_seq(lang, [["big", 9.0, 9.5], ["world", 10.01, 10.2]], start, end, n)
    assert ext2[0] == lang.get_word_index("world") and int((ext2 != 0).sum()) == 1
    # remove_word_timing: words that fall inside are spread evenly
    ext3 = data.extend_word_seq(lang, words, start, end, n, remove_word_timing=True)
    space = int(n / (4 + 1))
    assert [int(i) for i in torch.nonzero(ext3).flatten()] == [space, 2 * space, 3 * space, 4 * space]
    assert ext3[space] == lang.get_word_index("hello") and ext3[4 * space] == lang.get_word_index("world")


def test_words_to_tensor():
    lang = _lang()
    words = [["hello", 0.1, 0.2], ["zzz", 0.5, 0.6], ["world", 3.0, 3.1]]
    assert data.words_to_tensor(lang, words, 2.0).tolist() == [Vocab.SOS_token, lang.get_word_index("hello"), Vocab.UNK_token,
                                                             Vocab.EOS_token]
    assert data.words_to_tensor(lang, words).tolist()[-2] == lang.get_word_index("world")


def test_shard_round_trip_and_collate(tmp_path):
    rs = np.random.RandomState(0)
    lang = _lang()
    n_ext = 42                                             # the preprocessor stores 1.25 * n_poses frames
    samples = []
    for i in range(3):
        dur = n_ext / 15
        samples.append(([["hello", 5.0 + 0.2, 5.4], ["world", 5.0 + 1.5, 6.7]], rs.randn(n_ext, 43, 3).astype(np.float32),
                        rs.randn(n_ext, 42, 3).astype(np.float32), rs.randn(int(dur * 16000) - 100 * i).astype(np.float32),
                        (rs.rand(128, 86) * -80).astype(np.float16), {"vid": f"v{i}", "start_time": 5.0, "end_time": 5.0 + dur,
                                                                     "start_frame_no": np.int64(75 + i), "end_frame_no": np.int64(117 + i)}))   # numpy-typed aux
    data.write_shard(str(tmp_path / "shard0"), samples)
    ds = data.SpeechMotionShardDataset(str(tmp_path / "shard0"), 34, 15)
    ds.set_lang_model(lang)
    assert len(ds) == 3
    words_t, ext, pose, vec, audio, spec, aux = ds[1]
    assert pose.shape == (34, 129) and vec.shape == (34, 126) and audio.shape == (36267,) and spec.shape == (128, 70)
    assert torch.equal(vec, torch.from_numpy(samples[1][2][:34]).reshape(34, -1))
    assert torch.equal(spec, torch.from_numpy(samples[1][4][:, :70].astype(np.float32)))
    end = 5.0 + (n_ext / 15) * 34 / n_ext
    assert torch.equal(ext, data.extend_word_seq(lang, samples[1][0], 5.0, end, 34))
    assert words_t.tolist() == [Vocab.SOS_token, lang.get_word_index("hello"), lang.get_word_index("world"), Vocab.EOS_token]
    batch = data.default_collate_fn([ds[i] for i in range(3)])
    assert batch[2].shape == (3, 34) and batch[4].shape == (3, 34, 126) and batch[6].shape == (3, 128, 70)
    assert batch[7]["vid"] == ["v0", "v1", "v2"]
