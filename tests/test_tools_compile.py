"""The measurement scripts under tools/ and the repo-root entry points only ever run on the GPU box; a syntax
slip in one of them would surface there, minutes into a paid call.  Compile them all here."""
import glob
import os
import py_compile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_python_entry_points_and_tools_compile(tmp_path):
    files = sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + [
        os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    assert len(files) > 5
    for i, f in enumerate(files):
        py_compile.compile(f, cfile=str(tmp_path / f"{i}.pyc"), doraise=True)


def test_shell_tools_parse():
    import subprocess
    for f in sorted(glob.glob(os.path.join(ROOT, "tools", "*.sh"))):
        r = subprocess.run(["bash", "-n", f], capture_output=True, text=True)
        assert r.returncode == 0, (f, r.stderr)
