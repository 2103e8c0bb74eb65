"""Minimal stand-in for the reference's `concern.charsets.EnglishCharset` (concern/charsets.py:9-62,101-104):
38 classes = blank '\\t' (0), unknown '\\n' (1), then '0'-'9', 'A'-'Z' sorted.  Only what the hot-path modules use
(`len(charset)`, `.blank`, `.unknown`, label <-> string); the reference's own class is used when importable."""
import string

import numpy as np


class EnglishCharset:
    blank = 0
    unknown = 1
    blank_char = '\t'
    unknown_char = '\n'
    case_sensitive = False

    def __init__(self, **kwargs):
        chars = sorted(set(string.digits + string.ascii_uppercase))
        self._charset = [self.blank_char, self.unknown_char] + chars
        self._lut = {c: i for i, c in enumerate(self._charset)}

    def __len__(self):
        return len(self._charset)

    def __getitem__(self, index):
        return self._charset[index]

    def index(self, x):
        return self._lut.get(x if self.case_sensitive else x.upper(), self.unknown)

    def is_empty(self, index):
        return index == self.blank or index == self.unknown

    def string_to_label(self, string_input, max_size=32):
        target = np.zeros((max(max_size, len(string_input)),), dtype=np.int32)
        for i, c in enumerate(string_input):
            target[i] = self.index(c)
        return target

    def label_to_string(self, label):
        return "".join(self._charset[int(i)] for i in label if int(i) not in (self.unknown, self.blank))


def default_charset():
    try:
        from concern.charsets import DefaultCharset  # the reference's host code, when it is on sys.path
        return DefaultCharset()
    except Exception:
        return EnglishCharset()
