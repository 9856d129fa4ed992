"""Minimal ``Vocab`` type used as the speaker / language model handle (``z_obj``).

Same fields and default-token ids as the reference's scripts/model/vocab.py:8-38 (the hot path
only reads ``n_words`` and type-checks the object, hierarchy_net.py:76-80), but without importing
``fasttext`` at module import time.
"""


class Vocab:
    PAD_token, SOS_token, EOS_token, UNK_token = 0, 1, 2, 3

    def __init__(self, name, insert_default_tokens=True):
        self.name = name
        self.trimmed = False
        self.word_embedding_weights = None
        self.reset_dictionary(insert_default_tokens)

    def reset_dictionary(self, insert_default_tokens=True):
        self.word2index, self.word2count = {}, {}
        if insert_default_tokens:
            self.index2word = {0: "<PAD>", 1: "<SOS>", 2: "<EOS>", 3: "<UNK>"}
        else:
            self.index2word = {self.UNK_token: "<UNK>"}
        self.n_words = len(self.index2word)

    def index_word(self, word):
        if word in self.word2index:
            self.word2count[word] += 1
            return
        self.word2index[word] = self.n_words
        self.word2count[word] = 1
        self.index2word[self.n_words] = word
        self.n_words += 1

    def get_word_index(self, word):
        return self.word2index.get(word, self.UNK_token)


def is_vocab(obj) -> bool:
    """Duck-typed isinstance(obj, vocab.Vocab): accepts the reference's own Vocab objects too."""
    return obj is not None and hasattr(obj, "n_words") and hasattr(obj, "word2index")


def make_speaker_vocab(n_speakers: int) -> Vocab:
    """Speaker model as built by the data loader (insert_default_tokens=False => first id is 1)."""
    v = Vocab("vid", insert_default_tokens=False)
    for i in range(n_speakers):
        v.index_word(f"spk{i}")
    return v
