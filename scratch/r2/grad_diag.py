"""Per-tensor error of ONE update of the bench configuration (config 3, batch 512) in every tier."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
import numpy as np
import bench
from conftest import rel_err
from oracle import shinnosuke_oracle as O

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
X, Y = bench.synth('cfg3', 512)
np.random.seed(0)
spec, ishape = O.spec_cfg3()
net = O.OracleNet(spec, ishape)
net.params[-2] *= scale
p0 = [p.copy() for p in net.params]
net.fit(X.astype(np.float64), Y.astype(np.float64), 512, 1, shuffle=False)
names = []
for i in range(4):
    names += ['conv%d W' % (i + 1), 'conv%d b' % (i + 1), 'bn%d gamma' % (i + 1), 'bn%d beta' % (i + 1)]
names += ['dense W', 'dense b']
for prec in ['fp32', 'tf32x3', 'tf32', 'bf16']:
    m = bench.build_model('cfg3', prec)
    d = m.layers[-1]
    d.variables[0].output_tensor = np.asarray(d.variables[0].output_tensor) * scale
    m.fit(X, Y, batch_size=512, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    out = []
    for v, a, b, nm in zip(m.trainable_variables, p0, net.params, names):
        w = np.asarray(v.output_tensor).reshape(a.shape)
        out.append('%s %.1e' % (nm, rel_err(w - a, b - a)))
    print(prec, 'loss', m.train_loss[0], net.train_loss[0], '| update errors:', ', '.join(out), flush=True)
