"""The host / device shared texts under AddressSanitizer + UndefinedBehaviorSanitizer (CPU): vvenc_b200/csrc/rdoq_core.h (both engines of the fast RDOQ, the transform-skip and
BDPCM quantisers) with every check on, vvenc_b200/csrc/depquant_core.h without the signed-overflow check (its 64-bit distortion products wrap for extreme coefficients
exactly where DepQuant.cpp:652-668 wraps -- the device wraps by definition).  The kernels are thin wrappers around these texts, so an index that strays here would stray there.  Automatic variables are pattern-initialised
(-ftrivial-auto-var-init=pattern) and the golden results are still required: nothing depends on an uninitialised local."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _runtime(name):
    try:
        p = subprocess.run(['gcc', '-print-file-name=' + name], capture_output=True, text=True).stdout.strip()
    except OSError:
        return None
    return p if os.path.isabs(p) and os.path.exists(p) else None


@pytest.mark.parametrize("what,extra", [('rdoq', []), ('dq', ['-fno-sanitize=signed-integer-overflow'])])
def test_shared_texts_are_clean_under_asan_and_ubsan(tmp_path, what, extra):
    asan, ubsan = _runtime('libasan.so'), _runtime('libubsan.so')
    if not asan or not ubsan:
        pytest.skip('sanitizer runtimes not installed')
    lib = str(tmp_path / 'libshared_text_san.so')
    src = [os.path.join(ROOT, 'oracle', 'depquant_oracle.cpp'), os.path.join(ROOT, 'oracle', 'rdoq_oracle.cpp')]
    subprocess.check_call(['g++', '-O1', '-g', '-std=c++14', '-fPIC', '-shared', '-ffp-contract=off', '-fsanitize=address,undefined', '-fno-sanitize-recover=undefined', '-ftrivial-auto-var-init=pattern'] + extra + ['-o', lib] + src)
    env = dict(os.environ); env['LD_PRELOAD'] = asan + ' ' + ubsan; env['ASAN_OPTIONS'] = 'detect_leaks=0'
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', '_sanitizer_run.py'), lib, what], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and 'SANITIZER CLEAN' in out.stdout, (out.stdout[-500:], out.stderr[-3000:])
    assert int(out.stdout.split('SANITIZER CLEAN')[1].split()[0]) > (10000 if what == 'rdoq' else 200)
