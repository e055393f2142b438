import gzip
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "open-msspe-design_b200"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def zika_fasta() -> bytes:
    """96 pre-aligned Zika-like genomes: the reference's od-msspe/graph.fast, gzip-compressed data fixture."""
    with gzip.open(os.path.join(GOLDEN, "zika96_aligned.fa.gz"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as O
    O.build()
    return O
