"""The JNI glue (rafting_b200/csrc/jni/rafting_jni.c) cannot be built here (no JDK), but it must not rot: it is type-checked
with gcc against the real C-ABI headers and a stand-in jni.h, and every native method INTEGRATION.md's NativeEngine declares
must have its Java_..._NativeEngine_<name> definition in the glue."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "rafting_b200", "csrc", "jni", "rafting_jni.c")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_glue_type_checks_against_the_c_abi():
    res = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "jni_stub"),
                          "-I", os.path.join(ROOT, "include"), GLUE], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]


def test_every_native_method_of_the_integration_guide_is_defined():
    guide = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    declared = set(re.findall(r"static native [\w\[\]]+ (\w+)\(", guide))
    defined = set(re.findall(r"FN\((\w+)\)", open(GLUE).read()))
    assert len(declared) >= 30 and declared <= defined, sorted(declared - defined)
