#!/usr/bin/env python
"""Recipe that vendors the UNMODIFIED reference hot path next to the oracle (test / bench infrastructure, not product code).

    python oracle/build_ref.py            (build container: /root/reference is present)

copies the five modules SURVEY.md section 8(a) cites -- the Python sources where they lie under ``/root/reference/modules`` -- into
``oracle/_ref/modules/`` byte for byte (sha256 recorded in ``oracle/_ref/MANIFEST.json``).  ``oracle/_ref/`` is git-ignored (no
reference source enters the history) but travels to the GPU box with the snapshot, like the built ``.so``: there it is

  * the CPU arm of ``bench.py`` (``--impl reference`` / ``cpu_baseline.kind = "reference"``): the reference's own
    ``Gmm_nbit.estimate_from_y`` (modules/gmm_cplx_bussgang.py:166-243), single process and under the
    ``mp.Pool(cpu_count() // 2).starmap`` pattern of Bussgang_GMM.py:29-32, 282-287;
  * a second pin of the oracle (tests/test_oracle_golden.py::test_oracle_vs_vendored_reference).

Nothing under ``quantized_channel_estimation_b200/`` may import it (tests/test_host_cpu.py enforces that for ``oracle`` as a whole).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
FILES = ['modules/gmm_cplx_bussgang.py', 'modules/mofa_cplx_bussgang.py', 'modules/utils.py', 'modules/uniform_quantizer.py',
         'modules/lloyd_max_quantizer.py']


def build(reference_root='/root/reference', quiet=False):
    """Copy the cited modules; returns the destination, or None when the reference is not present (GPU box: prebuilt files are used)."""
    if not os.path.isdir(os.path.join(reference_root, 'modules')):
        return DST if os.path.exists(os.path.join(DST, 'MANIFEST.json')) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(reference_root, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, 'rb').read()).hexdigest()
    json.dump({'source': reference_root, 'files': manifest}, open(os.path.join(DST, 'MANIFEST.json'), 'w'), indent=1)
    if not quiet:
        print('vendored', len(FILES), 'reference modules into', DST)
    return DST


def available():
    return os.path.exists(os.path.join(DST, 'MANIFEST.json'))


def import_reference():
    """``(Gmm_nbit, Mofa, utils)`` of the vendored reference (``oracle/_ref`` first on sys.path: its modules import each other as
    ``modules.*``)."""
    if not available():
        raise RuntimeError('oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists')
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from modules.gmm_cplx_bussgang import Gmm_nbit
    from modules.mofa_cplx_bussgang import Mofa
    import modules.utils as ut
    return Gmm_nbit, Mofa, ut


if __name__ == '__main__':
    if build() is None:
        sys.exit('no reference tree and no prebuilt oracle/_ref')
