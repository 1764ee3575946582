"""CTCEncoder -- host-side text <-> label-id helper kept so that
``from pytorch_end2end import CTCLoss, CTCDecoder, CTCEncoder`` (pytorch_end2end/__init__.py:6)
keeps working.  No compute; behaviour follows pytorch_end2end/encoders/text_encoders.py:8-41."""
import numpy as np


class CTCEncoder:
    def __init__(self, characters, blank_id=0, transform_fn=str.upper):
        self.blank_id = blank_id
        self.transform_fn = transform_fn
        ids = (i for i in range(len(characters) + 1) if i != blank_id)
        self.char2id = {c: next(ids) for c in characters}
        self.id2char = {i: c for c, i in self.char2id.items()}
        self.id2char[blank_id] = ""
        self.num_symbols = len(self.id2char)

    def clean(self, text):
        return "".join(c for c in self.transform_fn(text) if c in self.char2id)

    def encode(self, text):
        return np.array([self.char2id[c] for c in self.clean(text)])

    def decode(self, ids_list):
        out, prev = [], None
        for i in ids_list:
            if i != prev and i != self.blank_id:
                out.append(self.id2char[i])
            prev = i
        return "".join(out)

    def decode_pure(self, ids_list):
        return "".join(self.id2char[i] for i in ids_list)
