"""Stage the UNMODIFIED reference (martinambrus/PitchExtractor) under ``baseline/_ref/`` so that it travels to the GPU
box (``baseline/_ref`` is git-ignored, not gpurun-ignored): ``bench.py --impl reference`` and the ``torch_gpu_baseline``
leg drive its own ``MelDataset._build_training_example`` / ``Collater`` / ``JDCNet`` / ``Trainer.run`` from there.

The reference has no ``setup.py`` / ``pyproject.toml`` (``pip install /root/reference`` does not apply): its "install" is
a byte-for-byte copy of the files the hot path imports.  Run by ``__graft_entry__.build()`` in the build container,
where ``/root/reference`` exists; a no-op elsewhere.  Nothing under ``pitchextractor_b200/`` imports these files.
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("model.py", "trainer.py", "optimizers.py", "meldataset.py", "f0_backends.py", "Utils/__init__.py",
         "Utils/synthetic.py", "Utils/dynamic_pitch_tools.py", "Configs/config.yml", "LICENSE")


def install(verbose=False):
    """-> DST if the reference is staged (now or earlier), else None."""
    if not os.path.isfile(os.path.join(SRC, "model.py")):
        return DST if os.path.isfile(os.path.join(DST, "model.py")) else None
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print("staged", rel)
    return DST


if __name__ == "__main__":
    print(install(verbose=True))
