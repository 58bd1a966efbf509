"""Copies the arrays of the reference's acceptance vectors
(xopto/mcml/test/reference/*.pkl: independent CUDA-MCML results used by
xopto/mcml/test/validate.py:219-866) that tests/test_gpu_validate.py needs into
tests/golden/validate_vectors.npz (this container only: /root/reference).

    python tests/golden/make_validate.py
"""
import os
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('XOPTO_REFERENCE', '/root/reference')
DATA_DIR = os.path.join(REF, 'xopto', 'mcml', 'test', 'reference')


def _load(name):
    with open(os.path.join(DATA_DIR, name + '.pkl'), 'rb') as f:
        return pickle.load(f)


def _layers(data):
    # rows of (d, n, mua, mus, g)
    return np.array([[L['basic']['d'], L['basic']['n'], L['basic']['mua'], L['basic']['mus'],
                      L['pf']['pfparams'][0]] for L in data['layers']], np.float64)


def main():
    out = {}
    d = _load('cuda_singlelayer_linesource_radial')
    out.update(line_layers=_layers(d), line_rmax=d['mc']['rmax'],
               line_axis=np.array([d['detector']['axis'][k] for k in ('start', 'stop', 'n')]),
               line_cosmin=d['detector']['cosmin'], line_reflectance=d['reflectance'])
    for key, name in (('single', 'cuda_singlelayer_uniformfiber'),
                      ('double', 'cuda_doublelayer_uniformfiber')):
        d = _load(name)
        lut = d['lut'] if isinstance(d['lut'], dict) else d['lut'][-1]
        out.update({
            key + '_layers': _layers(d), key + '_rmax': d['mc']['rmax'],
            key + '_fiber': np.array([d['source'][k] for k in ('dcore', 'dcladding', 'ncore', 'na')]),
            key + '_axis': np.array([d['detector']['axis'][k] for k in ('start', 'stop', 'n')]),
            key + '_cosmin': d['detector']['cosmin'],
            key + '_sds': np.asarray(d['detector']['fibers']['sds']),
            key + '_dcore': d['detector']['fibers']['dcore'],
            key + '_mua': lut['mua'], key + '_musr': lut['musr'],
            key + '_reflectance': d['reflectance']})
        if not isinstance(d['lut'], dict):
            out[key + '_top_mua_musr'] = np.array([d['lut'][0]['mua'][0], d['lut'][0]['musr'][0]])
    d = _load('single_layer_uniformfiber_trace')
    out.update(trace_layers=_layers(d), trace_rmax=d['mc']['rmax'],
               trace_fiber=np.array([d['source'][k] for k in ('dcore', 'dcladding', 'ncore', 'na')]),
               trace_axis=np.array([d['detector']['axis'][k] for k in ('start', 'stop', 'n')]),
               trace_cosmin=d['detector']['cosmin'], trace_maxlen=d['trace']['maxlen'],
               trace_mua=d['lut']['mua'], trace_musr=d['lut']['musr'])
    np.savez_compressed(os.path.join(HERE, 'validate_vectors.npz'), **out)
    for k, v in out.items():
        print(k, np.asarray(v).shape)


if __name__ == '__main__':
    main()
