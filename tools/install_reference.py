"""Copy the UNMODIFIED reference into baseline/_ref/ (git-ignored; it ships to the GPU box with the gpurun snapshot
exactly like the built .so) so that `bench.py --impl reference` times the reference's own code on the box's host
cores.  `pip install /root/reference` is not applicable: the reference has no setup.py / pyproject.toml.

    python tools/install_reference.py

Container-only (reads /root/reference).  Only bench.py's reference arm imports the copy, through the four import
shims of tools/ref_harness.py (stub pytorch_lightning / torchrl / matplotlib / seaborn, torch.load map_location,
zero placeholder y*.pt blobs); none of them touches hot-path arithmetic."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness  # noqa: E402


def install():
    if not os.path.isdir(ref_harness.REF_SRC):
        return None
    dst = os.path.abspath(ref_harness.INSTALLED)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    return ref_harness.copy_reference(dst)


if __name__ == "__main__":
    print(install())
