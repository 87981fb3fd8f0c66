"""The reference's entry points re-stated for the drop-in: a short training run writes a checkpoint that
inference.py loads."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_then_inference(tmp_path):
    wd = str(tmp_path)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--config", "config_energy", "--steps", "3",
                        "--n_graphs", "64", "--workdir", wd, "--name", "t", "--log_period", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ckpt = os.path.join(wd, "t", "model.pt")
    assert os.path.exists(ckpt)
    out = os.path.join(wd, "res.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "inference.py"), "--config", "config_energy", "--model_path", ckpt,
                        "--output_keys", "total_energy", "--n_graphs", "40", "--batch_size", "16", "--output_path", out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    z = np.load(out)
    assert z["total_energy"].shape == (40, 1) and np.isfinite(z["total_energy"]).all()


def test_train_energy_force(tmp_path):
    """force matching through the second-order mode (reference config_energy_force.py:30 loss)"""
    wd = str(tmp_path)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--config", "config_energy_force", "--steps", "3",
                        "--n_graphs", "64", "--workdir", wd, "--name", "f", "--log_period", "1"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.exists(os.path.join(wd, "f", "model.pt"))
    assert "step 2 loss" in r.stderr


def test_train_diffusion_score_matching(tmp_path):
    """reference train.py:70-215 (train_diffusion) on synthetic W4 molecules: VP-SDE loss, EMA, checkpoint"""
    wd = str(tmp_path)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--config", "config_diffusion", "--steps", "4",
                        "--workdir", wd, "--name", "d", "--log_period", "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.exists(os.path.join(wd, "d", "model.pt")) and "step 3 training_loss" in r.stderr
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--config", "config_diffusion_CA", "--steps", "3",
                        "--n_res", "150", "--workdir", wd, "--name", "p", "--log_period", "1"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "step 2 training_loss" in r.stderr
