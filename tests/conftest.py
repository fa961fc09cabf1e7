import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU and the built CUDA library")


@pytest.fixture(scope="session")
def oracle():
    import reface_oracle
    return reface_oracle


@pytest.fixture(scope="session")
def unet_sd(oracle):
    return oracle.init_state_dict(oracle.unet_spec(), 0)


@pytest.fixture(scope="session")
def vae_sd(oracle):
    return oracle.init_state_dict(oracle.vae_spec(), 0)


@pytest.fixture(scope="session")
def clip_sd(oracle):
    return oracle.init_state_dict(oracle.clip_spec(), 0)


@pytest.fixture(scope="session")
def arc_sd(oracle):
    return oracle.init_state_dict(oracle.arcface_spec(), 0)


@pytest.fixture(scope="session")
def fusion_sd(oracle):
    return oracle.init_state_dict(oracle.fusion_spec(), 0)


@pytest.fixture(scope="session")
def engine():
    import torch
    from reface_b200.runtime import Engine
    assert torch.cuda.is_available(), "gpu tests need a GPU"
    eng = Engine(0, arena_bytes=24 << 30)
    for k, v in os.environ.items():          # e.g. RFB_GEMM_PAIR=0 to A/B a kernel generation
        if k.startswith("RFB_") and k != "RFB_CPU_THREADS":
            eng.set_option(k[4:].lower(), int(v))
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def unet_engine(engine, unet_sd):
    engine.load_state_dict(unet_sd)
    engine.build_unet()
    return engine
