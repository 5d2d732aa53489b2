"""Minimal stand-in for the `termcolor` package imported by /root/reference/seq_lattice/models.py:4."""


def colored(text, *args, **kwargs):
    return text
