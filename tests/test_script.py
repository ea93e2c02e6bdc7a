"""dbat_b200/script.py (`rundbatscript`) on the three script projects the reference ships, with the oracle
as numerical backend on CPU and the device in the `-m gpu` twin: the XML is parsed, the struct built and the
operations run by the package; the result file must be the reference's."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest

from dbat_b200.script import rundbatscript
from test_report_golden import GOLD, report_diff


def oracle_backend():
    from oracle.photogrammetry import resect, forwintersect
    from oracle.bundle import bundle, bundle_cov
    return NS(resect=resect, forwintersect=forwintersect, bundle=bundle, bundle_cov=bundle_cov)


@pytest.mark.parametrize('name,rtol', [('camcaldemo', 0.0), ('sxb', 1e-4), ('romabundledemo', 0.0)])
def test_script_projects_reproduce_the_reference_reports(name, rtol):
    root = os.path.join(GOLD, name)
    s, E, lines = rundbatscript(os.path.join(root, name + '.xml'), backend=oracle_backend(), write=False)
    assert E.code == 0
    d = report_diff(lines, os.path.join(root, 'result', 'report.txt'))
    if rtol:                              # sxb: start values from a resection at 1e6 m (see test_report_golden)
        assert len(d) == 1 and 'First error' in d[0][1]
        d = report_diff(lines, os.path.join(root, 'result', 'report.txt'), rtol=rtol)
    assert d == []


def test_script_writes_report_and_camera_files(tmp_path):
    """Output section: the report goes to the path the script names, the calibrated camera to the io file
    in the user's sign conventions with 18 significant digits (result/c4040z.xml of the reference)."""
    import re
    import shutil
    root = tmp_path / 'camcaldemo'
    shutil.copytree(os.path.join(GOLD, 'camcaldemo'), root)
    shutil.rmtree(root / 'result')
    os.makedirs(root / 'result')
    s, E, lines = rundbatscript(str(root / 'camcaldemo.xml'), backend=oracle_backend())
    assert open(root / 'result' / 'report.txt').read().split('\n')[:-1] == lines
    got = open(root / 'result' / 'c4040z.xml').read()
    ref = open(os.path.join(GOLD, 'camcaldemo', 'result', 'c4040z.xml')).read()
    num = lambda tag, txt: [float(v) for v in re.search(r'<%s>([^<]*)</%s>' % (tag, tag), txt).group(1).split(',')]
    for tag in ('cc', 'pp', 'K', 'P', 'aspect', 'skew', 'sensor', 'image'):
        np.testing.assert_allclose(num(tag, got), num(tag, ref), rtol=1e-9 if tag != 'sensor' else 1e-5, atol=1e-14,
                                   err_msg=tag)
    # eo file: 18-digit EO values and deviations; residual file: the 50 largest image residuals
    body = lambda path: [l.rstrip('\n') for l in open(path) if not l.startswith('#')]
    got, ref = body(root / 'result' / 'camera_stations.txt'), body(os.path.join(GOLD, 'camcaldemo', 'result', 'camera_stations.txt'))
    assert len(got) == len(ref) == 21
    for a, b in zip(got, ref):
        ta, tb = a.split(', '), b.split(', ')
        assert ta[:2] == tb[:2] and ta[-1] == tb[-1]
        np.testing.assert_allclose([float(v) for v in ta[2:8]], [float(v) for v in tb[2:8]], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose([float(v) for v in ta[8:14]], [float(v) for v in tb[8:14]], rtol=1e-6)
    got, ref = body(root / 'result' / 'top_residuals.txt'), body(os.path.join(GOLD, 'camcaldemo', 'result', 'top_residuals.txt'))
    assert len(got) == len(ref) == 50
    for a, b in zip(got, ref):
        ta, tb = a.split(', '), b.split(', ')
        assert ta[:4] == tb[:4]
        np.testing.assert_allclose([float(v) for v in ta[4:]], [float(v) for v in tb[4:]], rtol=2e-5, atol=2e-6)


def test_script_errors_are_loud(tmp_path):
    bad = tmp_path / 'bad.xml'
    bad.write_text('<document dbat_script_version="2.0"><input/><operations/><output/></document>')
    with pytest.raises(ValueError):
        rundbatscript(str(bad), backend=oracle_backend())
    src = open(os.path.join(GOLD, 'camcaldemo', 'camcaldemo.xml')).read()
    import shutil
    cases = ([('<operation>spatial_resection</operation>', '<operation>levitate</operation>')],   # unknown operation
             [('min_rays="2"', 'min_rays="30"')],                                                  # ray count check
             [('<image_pts>', '<image_points>'), ('</image_pts>', '</image_points>')],             # unknown / missing field
             [('format="id,label,x,y,z"', 'format="id,label,x,y"')])                               # table / format mismatch
    for n, edits in enumerate(cases):
        root = tmp_path / ('case%d' % n)
        shutil.copytree(os.path.join(GOLD, 'camcaldemo'), root)
        txt = src
        for old, new in edits:
            assert old in txt
            txt = txt.replace(old, new)
        (root / 'camcaldemo.xml').write_text(txt)
        with pytest.raises(ValueError):
            rundbatscript(str(root / 'camcaldemo.xml'), backend=oracle_backend(), write=False)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['camcaldemo', 'sxb', 'romabundledemo'])
def test_script_projects_on_the_device(name):
    """`python -m dbat_b200.script project.xml` as a user runs it: everything numerical on the device."""
    root = os.path.join(GOLD, name)
    s, E, lines = rundbatscript(os.path.join(root, name + '.xml'), write=False)
    assert E.code == 0
    assert report_diff(lines, os.path.join(root, 'result', 'report.txt'), rtol=1e-5, first_error_rtol=1e-3) == []


def test_script_initial_value_fields(tmp_path):
    """set_initial_values/io field by field (parsesetinitialiovalues.m): with cc = 7.3 written out, default
    principal point and zero distortion, the camcal script becomes the PhotoModeler demo `camcaldemo.m`
    (EXIF focal length 7.3 instead of the script's 7.5) up to the sensor width - the script derives it from the
    pixel size ('auto'), the export states 7.25319 mm, which moves the default principal point by 1.5 um - and
    lands on that demo's numbers: 9 iterations instead of the script's 8, first error 30873.9 to 0.05 %,
    sigma0 1.6148."""
    import shutil
    root = tmp_path / 'camcaldemo'
    shutil.copytree(os.path.join(GOLD, 'camcaldemo'), root)
    src = open(root / 'camcaldemo.xml').read()
    old = '<io>\n          <all>default</all>\n        </io>'
    assert old in src
    new = ('<io><all>default</all><cc>7.3</cc><pp>default</pp><K>0,0,0</K><P>default</P>'
           '<aspect>1</aspect><skew>0</skew></io>')
    (root / 'camcaldemo.xml').write_text(src.replace(old, new))
    s, E, lines = rundbatscript(str(root / 'camcaldemo.xml'), backend=oracle_backend(), write=False)
    assert E.code == 0 and E.usedIters == 9
    assert abs(E.res[0] / 30873.9 - 1) < 5e-4 and abs(E.s0 - 1.6148) < 6e-5
