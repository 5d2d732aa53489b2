"""Stand-in for `termcolor` (seq_lattice/models.py:4) when the real package is absent."""


def colored(text, *args, **kwargs):
    return text
