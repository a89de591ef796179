"""Python handle on the CPU kernel-body emulator (tests/emu/libblingemu.so). TEST INFRASTRUCTURE ONLY."""
import subprocess
from pathlib import Path

from bling_b200.api import Context

_DIR = Path(__file__).resolve().parent


def build():
    subprocess.check_call(["make", "-s", "-C", str(_DIR)])
    return _DIR / "libblingemu.so"


class EmuContext(Context):
    _lib_path = _DIR / "libblingemu.so"
    _prefix = "blingemu"

    def __init__(self):
        build()
        super().__init__(0)
