"""CPU, only where the reference tree is mounted (/root/reference; not on the GPU box): every shipped fits/*.fit file
imports to the same keys and the same BSDF strings as the unmodified reference's io::importFIT (SURVEY.md fact 10:
bagher_sgd.fit overflows float and fails on both sides)."""
import glob
import os

import pytest

FITS = sorted(glob.glob("/root/reference/fits/*.fit"))


@pytest.mark.skipif(not FITS, reason="reference tree not mounted")
def test_all_shipped_fit_files_import_like_the_reference(ref):
    import bbm_b200 as bb
    assert len(FITS) == 14
    imported = 0
    for path in FITS:
        try:
            want = ref.import_fit(path)
        except Exception:
            want = None
        if want is None:
            assert os.path.basename(path) == "bagher_sgd.fit"          # c = [1.31522e+49, ...] overflows std::stof
            with pytest.raises(bb.BbmError):
                bb.import_fit(path)
            continue
        got = bb.import_fit(path)
        assert list(got) == list(want), path
        for k in want:
            # forward parameter order on our side, per-lobe reversed in the reference's run-time aggregate (fact 14):
            # the STRINGS agree, which is what the file format is
            assert got[k].to_string() == want[k], (path, k)
        imported += 1
    assert imported == 13
