"""Developer diagnostic: runs each stage of the CUDA path against the oracle and prints differences."""
import json, os, sys, time
import numpy as np, cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nonlocal_image_edit_b200 as nb
from oracle import nle_oracle as O
G = os.path.join(ROOT, 'tests/golden')
man = {m['name']: m for m in json.load(open(f'{G}/manifest.json'))}

def t(msg, fn):
    t0 = time.time()
    try:
        r = fn(); print(f'[ok] {msg} ({time.time()-t0:.2f}s)', flush=True); return r
    except Exception as e:
        print(f'[FAIL] {msg}: {type(e).__name__}: {e}', flush=True); return None

# 1. sampling
def samp():
    for (r, c, a, b) in [(736, 491, 20, 10), (100, 100, 40, 40), (7, 5, 7, 5), (267, 400, 10, 20)]:
        s, rest = nb.sampleIndices(r, c, a, b)
        so, ro = O.sample_pixels(r, c, a, b)
        assert np.array_equal(s, so) and np.array_equal(rest, ro), (r, c, a, b)
t('sampling bit-exact', samp)

# 2. eig
def eig():
    R = np.array([[2., -1, 0], [-1, 2, -1], [0, -1, 2]])
    U, D = nb.eigenDecomposition(R)
    print('   D', D, 'recon', np.abs((U * D) @ U.T - R).max(), 'orth', np.abs(U.T @ U - np.eye(3)).max())
    rng = np.random.default_rng(0)
    for n in (5, 16, 33, 100, 257):
        A = rng.standard_normal((n, n)); A = (A + A.T) / 2
        t0 = time.time(); U, D = nb.eigenDecomposition(A, eps=-1e300); dt = time.time() - t0
        w = np.linalg.eigvalsh(A)[::-1]
        print(f'   n={n} eig err {np.abs(D - w).max():.2e} orth {np.abs(U.T @ U - np.eye(n)).max():.2e} resid {np.abs(A @ U - U * D).max():.2e} {dt*1e3:.1f} ms')
t('eig', eig)

# 3. pipeline
def pipe(name, crop=None):
    m = man[name]; img = cv2.imread(f"{G}/{name}_input.png")
    if crop: img = np.ascontiguousarray(img[:crop[0], :crop[1]])
    a = (m['n_row_samples'], m['n_col_samples'], m['hx'], m['hy'], m['n_sinkhorn_iter'], m['n_eigen_vectors'])
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2Lab); L0 = np.ascontiguousarray(lab[:, :, 0])
    f = nb.NLEFilter()
    t0 = time.time(); f.trainFilter(L0, *a); dt = time.time() - t0
    inf = f.info()
    print(f'   {name} p={inf.p} r={inf.r} r2={inf.r2} k={inf.k} sweeps={list(inf.eig_sweeps)} train {dt*1e3:.1f} ms; stage ms {np.round(f.stage(8), 2)}')
    fo = O.train_dense(L0.astype(np.float64), *a)
    st = fo.stages
    print(f'   oracle r={st["r"]} r2={st["r2"]} k={fo.eigvals.size}')
    Ka = f.stage(0).reshape(inf.p, inf.p, order='F'); print('   Ka diff', np.abs(Ka - st['Ka']).max())
    lam = f.stage(1); print('   lam rel diff', (np.abs(lam - st['lam'][:lam.size]) / st['lam'][:lam.size]).max() if lam.size == st['lam'].size else 'size mismatch')
    if inf.r == st['r']:
        cg = f.stage(3); co = np.empty_like(cg); co[st['perm']] = st['c']
        print('   c rel diff', np.abs(cg - co).max() / np.abs(co).max())
        print('   rvec rel diff', np.abs(f.stage(2) - st['rvec_head']).max() / np.abs(st['rvec_head']).max())
        Wa = f.stage(4).reshape(inf.r, inf.r, order='F'); print('   Wa diff', np.abs(Wa - st['Wa']).max(), 'scale', np.abs(st['Wa']).max())
        Q = f.stage(5).reshape(inf.r, inf.r, order='F'); print('   Q diff', np.abs(Q - st['Q']).max(), 'scale', np.abs(st['Q']).max())
    S = f.eigvals; kq = min(S.size, fo.eigvals.size)
    print('   Sq', S[:4], 'rel diff', (np.abs(S[:kq] - fo.eigvals[:kq]) / fo.eigvals[:kq]).max())
    Lg = f.enhanceLuminance(L0, m['weights']); Lo = O.enhance_luminance(fo, L0, m['weights'])
    d = np.abs(Lg.astype(int) - Lo.astype(int)); print(f'   L out: max {d.max()} eq {(d == 0).mean():.5f} le1 {(d <= 1).mean():.5f}')
    if crop is None:
        gold = cv2.imread(f"{G}/{name}_golden.png"); out = f.enhance(img, m['weights'])
        dg = np.abs(out.astype(int) - gold.astype(int)); print(f'   vs golden: max {dg.max()} le1 {(dg <= 1).mean():.5f}')
names = sys.argv[1:] or ['flower', 'forest', 'brickwall']
t('pipeline crop', lambda: pipe('flower', crop=(120, 160)))
for n in names:
    t('pipeline ' + n, lambda: pipe(n))
print('launches', nb.load().nle_b200_launch_count(0))
