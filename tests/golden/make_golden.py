#!/usr/bin/env python
"""Regenerate tests/golden/ from the read-only reference checkout (/root/reference).

Run in the build container only (the GPU box has no /root/reference; tests read the
committed copies).  Copies DATA fixtures (inputs + the reference's golden result
files) — never reference source code.
"""
import os
import shutil
import sys

REF = os.environ.get('DBAT_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = {
    'camcaldemo': [
        'data/script/camcaldemo/camcaldemo.xml',
        'data/script/camcaldemo/images/images.txt',
        'data/script/camcaldemo/measurements/markpts.txt',
        'data/script/camcaldemo/reference/camcal-fixed.txt',
        'data/script/camcaldemo/result/c4040z.xml',
        'data/script/camcaldemo/result/camera_stations.txt',
        'data/script/camcaldemo/result/report.txt',
        'data/script/camcaldemo/result/top_residuals.txt',
    ],
    'sxb': [
        'data/script/sxb/sxb.xml',
        'data/script/sxb/images/images.txt',
        'data/script/sxb/measurements/markpts.txt',
        'data/script/sxb/measurements/smartpts.txt',
        'data/script/sxb/reference/sxb-control.txt',
        'data/script/sxb/result/report.txt',
    ],
    'romabundledemo': [                                     # markpts.txt (2.6 MB) is stored xz-compressed
        'data/script/romabundledemo/romabundledemo.xml',
        'data/script/romabundledemo/cameras/EOS5DMarkII.xml',
        'data/script/romabundledemo/images/images.txt',
        'data/script/romabundledemo/prior/initial_eo.txt',
        'data/script/romabundledemo/measurements/markpts.txt',
        'data/script/romabundledemo/result/report.txt',
        'data/script/romabundledemo/result/EOS5DMarkII.xml',
    ],
    'prague2016cam': [
        'data/prague2016/cam/pmexports/weighted-no-orient-pmexport.txt',
        'data/prague2016/cam/pmexports/fixed-no-orient-pmexport.txt',
        'data/prague2016/cam/ref/ctrlpts-weighted.txt',
        'data/prague2016/cam/ref/ctrlpts-fixed.txt',
        'data/prague2016/cam/dbatexports/weighted-no-orient-dbatreport.txt',
        'data/prague2016/cam/dbatexports/fixed-no-orient-dbatreport.txt',
        'data/prague2016/cam/pmexports/weighted-with-orient-pmexport.txt',
        'data/prague2016/cam/pmexports/fixed-with-orient-pmexport.txt',
        'data/prague2016/cam/dbatexports/weighted-with-orient-dbatreport.txt',
        'data/prague2016/cam/pmexports/weighted-no-orient-3dpts.txt',      # PhotoModeler's own result tables
        'data/prague2016/cam/pmexports/fixed-no-orient-3dpts.txt',
        'data/prague2016/cam/pmexports/weighted-no-orient-pmreport.txt',
        'data/prague2016/cam/pmexports/fixed-no-orient-pmreport.txt',
        'data/prague2016/cam/dbatexports/fixed-with-orient-dbatreport.txt',
    ],
    'stpierre': [
        'data/hamburg2017/stpierre/pmexports/C5_reduced-pmexport.txt',
    ],
    'camcalpm': [
        'data/dbat/pmexports/camcal-pmexport.txt',
        'data/dbat/pmexports/camcal-pmexport-1ray.txt',
        'data/dbat/pmexports/camcal-pmexport-missing-obs.txt',
        'data/dbat/pmexports/camcal-pmexport5.txt',
        'data/dbat/ref/camcal-fixed.txt',
        'data/dbat/dbatexports/camcal-dbatreport-1ray.txt',
        'data/dbat/dbatexports/camcal-dbatreport-missing-obs.txt',
        'data/dbat/dbatexports/camcal-dbatreport-no-datum.txt',
        'data/dbat/dbatexports/camcal-dbatreport5.txt',
    ],
    'prague2016sxb': [
        'data/prague2016/sxb/pmexports/f-op0-no-orient-pmexport.txt',
        'data/prague2016/sxb/pmexports/w-op0-no-orient-pmexport.txt',
        'data/prague2016/sxb/pmexports/w-op1-no-orient-pmexport.txt',
        'data/prague2016/sxb/pmexports/wsmart-no-orient-pmexport.txt',
        'data/prague2016/sxb/pmexports/wsmart-with-orient-pmexport.txt',
        'data/prague2016/sxb/ref/fake-camera-positions.txt',
        'data/prague2016/sxb/dbatexports/sxb-prior-eo-dbatreport.txt',
        'data/prague2016/sxb/dbatexports/sxb-no-prior-eo-dbatreport.txt',
        'data/prague2016/sxb/ref/ctrlpts-fixed.txt',
        'data/prague2016/sxb/ref/ctrlpts-weighted.txt',
        'data/prague2016/sxb/dbatexports/f-op0-no-orient-dbatreport.txt',
        'data/prague2016/sxb/dbatexports/w-op0-no-orient-dbatreport.txt',
        'data/prague2016/sxb/dbatexports/w-op1-no-orient-dbatreport.txt',
        'data/prague2016/sxb/dbatexports/wsmart-no-orient-dbatreport.txt',
        'data/prague2016/sxb/pmexports/f-op0-with-orient-pmexport.txt',
        'data/prague2016/sxb/pmexports/w-op0-with-orient-pmexport.txt',
        'data/prague2016/sxb/pmexports/w-op1-with-orient-pmexport.txt',
        'data/prague2016/sxb/dbatexports/f-op0-with-orient-dbatreport.txt',
        'data/prague2016/sxb/dbatexports/w-op0-with-orient-dbatreport.txt',
        'data/prague2016/sxb/dbatexports/w-op1-with-orient-dbatreport.txt',
        'data/prague2016/sxb/dbatexports/wsmart-with-orient-dbatreport.txt',
        'data/prague2016/sxb/psprojects/sxb.psz',                       # PhotoScan archive (XML + PLY), 150 kB
        'data/prague2016/sxb/psprojects/sxb-dbatreport.txt',
        'data/prague2016/sxb/dbatexports/sxb-dbatreport.txt',
        'data/prague2016/sxb/psprojects/sxb-psstats-prefilt.txt',
        'data/prague2016/sxb/ref/ctrlpts-weighted-raw.txt',
    ],
    'dbatexports': [
        'data/dbat/dbatexports/camcal-dbatreport.txt',
        'data/dbat/dbatexports/camcal-dbatreport-model2.txt',
        'data/dbat/dbatexports/camcal-dbatreport-model3.txt',
        'data/dbat/dbatexports/camcal-dbatreport-model4.txt',
        'data/dbat/dbatexports/camcal-dbatreport-model5.txt',
        'data/dbat/dbatexports/camcal-dbatreport-model1.txt',
        'data/dbat/dbatexports/camcal-dbatreport-model-1.txt',
    ],
}


def main():
    if not os.path.isdir(REF):
        sys.exit('reference checkout not found at %s' % REF)
    for sub, files in FILES.items():
        for f in files:
            rel = f.split(sub + '/', 1)[1] if sub + '/' in f else os.path.basename(f)
            if sub == 'prague2016cam':
                rel = f.split('data/prague2016/cam/', 1)[1]
            if sub == 'prague2016sxb':
                rel = f.split('data/prague2016/sxb/', 1)[1]
            if sub in ('stpierre', 'camcalpm'):
                rel = os.path.basename(f)
            dst = os.path.join(HERE, sub, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if os.path.getsize(os.path.join(REF, f)) > (1 << 20):
                import lzma
                dst += '.xz'
                with open(os.path.join(REF, f), 'rb') as src, lzma.open(dst, 'wb', preset=9) as out:
                    shutil.copyfileobj(src, out)
            else:
                shutil.copyfile(os.path.join(REF, f), dst)
            print('copied', f, '->', os.path.relpath(dst, HERE))


if __name__ == '__main__':
    main()
