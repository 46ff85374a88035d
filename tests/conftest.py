"""pytest configuration: markers and import paths.

The helper modules (qd_cases) live beside the tests and are imported by bare name,
because an unrelated installed package called ``tests`` shadows ``tests.*`` imports
in this image.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
