"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz|json by running the UNMODIFIED reference
(oracle/_ref/libbbmref_{float,double}.so, built by oracle/Makefile from /root/reference) on fixed
seeded inputs.  The reference ships no known-answer vectors; these files pin
the oracle.  Re-run with:  make -C oracle ref && python -m oracle.gen_golden

Key of the fixtures (bin indices depend on it): g++ 13.3.0, glibc 2.39,
-O3 -DNDEBUG -ffp-contract=off, x86-64, reference bbm 0.5.1.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.refbind import METRICS, Ref, sph_desc  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF_FITS = "/root/reference/fits"

MODELS = ("Lambertian OrenNayar Phong NganBlinnPhong Lafortune NganLafortune Ward WardDuer WardDuerGeislerMoroder NganWard "
          "NganWardDuer AshikhminShirley AshikhminShirleyFull NganAshikhminShirley LowAshikhminShirley CookTorrance LowCookTorrance "
          "NganCookTorrance CookTorranceWalter CookTorranceHeitz GGX GGXHeitz PhongWalter LowMicrofacet LowMicrofacetFit LowSmooth "
          "Ribardiere RibardiereAnisotropic Bagher EPD He HeWestin HeHolzschuch NganHe").split()

# non-default parameter sets: sharp / rough / anisotropic cases, plus entries of the shipped fits
EXTRA = {
    "OrenNayar": ["OrenNayar([0.3, 0.6, 0.2], 0.7)"],
    "Phong": ["Phong([0.2, 0.3, 0.4], 800)"],
    "Lafortune": ["Lafortune([0.3, 0.2, 0.1], [-0.9, -0.7], 0.8, 50)"],
    "NganLafortune": ["NganLafortune([0.3, 0.2, 0.1], -0.6, 0.55, 20)"],
    "Ward": ["Ward([0.3, 0.2, 0.1], [0.05, 0.3])"],
    "WardDuer": ["WardDuer([0.3, 0.2, 0.1], [0.2, 0.04])"],
    "WardDuerGeislerMoroder": ["WardDuerGeislerMoroder([0.3, 0.2, 0.1], [0.07, 0.15])"],
    "AshikhminShirley": ["AshikhminShirley([0.05, 0.3, 0.6], [10, 900])"],
    "AshikhminShirleyFull": ["AshikhminShirleyFull([0.4, 0.2, 0.1], [0.05, 0.08, 0.1], [300, 20])"],
    "CookTorrance": ["CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5)", "CookTorrance([0.1, 0.2, 0.3], 0.02, 2.5)"],
    "NganCookTorrance": ["NganCookTorrance([0.1, 0.2, 0.3], 0.05, 0.3)"],
    "CookTorranceHeitz": ["CookTorranceHeitz([0.1, 0.2, 0.3], [0.05, 0.4], 1.8)"],
    "GGX": ["GGX([0.1, 0.2, 0.3], 0.01, 1.5)", "GGX([0.1, 0.2, 0.3], 0.5, 1.5)"],
    "GGXHeitz": ["GGXHeitz([0.1, 0.2, 0.3], [0.3, 0.03], 2.0)"],
    "PhongWalter": ["PhongWalter([0.1, 0.2, 0.3], 500, 1.6)"],
    "LowMicrofacet": ["LowMicrofacet([2, 3, 4], 5000, 1.4, 1.6)"],
    "LowSmooth": ["LowSmooth([2, 3, 4], 300, 2.2, 1.5)"],
    "Ribardiere": ["Ribardiere([0.1, 0.2, 0.3], 0.05, 1.8, 1.5)"],
    "RibardiereAnisotropic": ["RibardiereAnisotropic([0.1, 0.2, 0.3], [0.05, 0.3], 4.0, 1.5)"],
    "Bagher": ["Bagher(albedo = [0.2, 0.3, 0.4], alpha = [0.02, 0.05, 0.3], p = [0.3, 0.9, 1.5], eta = [[0.9, 0.8, 0.7], [0.1, -0.2, 0.3]], "
               "K = [3, 7, 12], Lambda = [0.5, 1.5, 2], c = [0.8, 1.2, 2], theta0 = [1.2, 1.0, 0.8], k = [1.5, 1, 0.7])"],
    "EPD": ["EPD(0.05, 0.5, [1.5, 0.5])", "EPD(0.3, 2.0, [0.2, 3.0])", "EPD(0.003, 1.0, [1.3, 0])", "EPD(0.5, 5.0, [2.0, 1.0])"],
    "He": ["He(0.05, 1.5, [[0.2, 0.9, 1.4], [3.0, 2.4, 1.9]])"],
    "HeWestin": ["HeWestin(0.0318, 0.3, [[0.2, 0.9, 1.4], [3.0, 2.4, 1.9]])"],
    "HeHolzschuch": ["HeHolzschuch(0.4, 6.0, [[1.5, 1.5, 1.5], [0, 0, 0]])"],
    "NganHe": ["NganHe([0.1, 0.2, 0.3], 0.3, 24, 1.5)"],
}
FIT_PICKS = {"ngan_cooktorrance.fit": ["alum-bronze", "blue-metallic-paint"], "low_cooktorrance_E1.fit": ["aluminium", "nylon"],
             "ngan_ward.fit": ["alum-bronze"], "ngan_wardduer.fit": ["nickel"], "ngan_blinnphong.fit": ["alum-bronze"],
             "ngan_ashikhminshirley.fit": ["alum-bronze"], "low_ashikhminshirley_E2.fit": ["gold-paint"],
             "low_lowmicrofacet_E2.fit": ["chrome"], "low_lowsmooth_E2.fit": ["chrome"], "ngan_he.fit": ["alum-bronze"],
             "ngan_lafortune.fit": ["alum-bronze", "violet-acrylic"]}


def directions(rng, n):
    """upper hemisphere + a few below-horizon, exactly-grazing and normal-incidence cases"""
    z = rng.random(n)
    ph = rng.random(n) * 2 * np.pi
    s = np.sqrt(1 - z * z)
    d = np.stack([s * np.cos(ph), s * np.sin(ph), z], 1).astype(np.float32)
    d[0] = (0, 0, 1)
    d[1] = (1, 0, 0)
    d[2] = (0.6, 0.0, -0.8)
    d[3] = (np.float32(np.sqrt(0.5)), np.float32(np.sqrt(0.5)), 0)
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    ref, refd = Ref("float"), Ref("double")
    rng = np.random.default_rng(20261018)
    n = 192
    inn, out = directions(rng, n), directions(rng, n)[::-1].copy()
    xi = rng.random((n, 2)).astype(np.float32)
    xi[0] = (0, 0); xi[1] = (1, 1); xi[2] = (0.5, 1.5); xi[3] = (0.25, 0.75)
    models = {}
    fits = {}
    for f, keys in FIT_PICKS.items():
        entries = ref.import_fit(os.path.join(REF_FITS, f))
        for k in keys:
            fits[f + ":" + k] = entries[k]
    cases = []
    for m in MODELS:
        cases.append(m + "()")
        cases += EXTRA.get(m, [])
    cases += list(fits.values())
    cases += ["Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))",
              "Aggregate(Lambertian([0.1, 0.1, 0.1]), GGX([0.5, 0.4, 0.3], 0.15, 1.7), Phong([0.2, 0.2, 0.2], 60))",
              "Aggregate(GGX())"]
    arrays = {"in": inn, "out": out, "xi": xi}
    for ci, s in enumerate(cases):
        rec = {"string": s, "canonical": ref.to_string(s)}
        for w, name in ((0, "values"), (1, "default"), (2, "lower"), (3, "upper")):
            rec[name] = ref.params(s, w).tolist()
        models[f"case{ci}"] = rec
        for comp in (3, 1, 2):
            arrays[f"case{ci}_eval_c{comp}"] = ref.eval(s, inn, out, comp)
            arrays[f"case{ci}_pdf_c{comp}"] = ref.pdf(s, inn, out, comp)
            arrays[f"case{ci}_refl_c{comp}"] = ref.reflectance(s, out, comp)
            d, p, fl = ref.sample(s, out, xi, comp)
            arrays[f"case{ci}_sdir_c{comp}"] = d
            arrays[f"case{ci}_spdf_c{comp}"] = p
            arrays[f"case{ci}_sflag_c{comp}"] = fl.astype(np.int8)
    np.savez_compressed(os.path.join(OUT, "models.npz"), **arrays)
    json.dump({"cases": models, "fits": fits}, open(os.path.join(OUT, "models.json"), "w"), indent=1)

    # per-model attribute layout from the reference's own reflection (flag bits probed one at a time)
    layout = {}
    for m in MODELS:
        layout[m] = {"string": ref.to_string(m + "()"),
                     "flags": {str(bit): {"default": ref.params(m + "()", 1, bit).tolist(), "lower": ref.params(m + "()", 2, bit).tolist(),
                                          "upper": ref.params(m + "()", 3, bit).tolist()} for bit in (1, 2, 4, 8, 16)}}
    json.dump(layout, open(os.path.join(OUT, "model_layout.json"), "w"), indent=1)

    # linearizers
    lin = {}
    m = 4096
    a, b = directions(rng, m), directions(rng, m)
    lin["pairs_in"], lin["pairs_out"] = a, b
    lin["pairs_index"] = ref.merl_index(a, b).astype(np.uint32)
    idx = np.sort(rng.choice(1458000, 2048, replace=False)).astype(np.uint32)
    idx[:4] = (0, 1, 179, 180); idx[-1] = 1457999
    gi = np.empty((len(idx), 3), np.float32); go = np.empty((len(idx), 3), np.float32)
    for k, i in enumerate(idx):
        x, y = ref.merl_dirs(int(i), 1)
        gi[k], go[k] = x[0], y[0]
    lin["grid_idx"], lin["grid_in"], lin["grid_out"] = idx, gi, go
    lin["grid_index_of_dirs"] = ref.merl_index(gi, go).astype(np.uint32)
    d = sph_desc((12, 7), (5, 6))
    lin["sph_in"], lin["sph_out"] = ref.spherical_dirs(d, 0, 12 * 7 * 5 * 6)
    d2 = sph_desc((9, 4), (1, 5), start_in=(0.1, 0.05), end_in=(3.0, 1.4), start_out=(0.0, 0.2), end_out=(6.0, 1.5))
    lin["sph2_in"], lin["sph2_out"] = ref.spherical_dirs(d2, 0, 9 * 4 * 1 * 5)
    np.savez_compressed(os.path.join(OUT, "linearizers.npz"), **lin)

    # losses: per-sample terms (floatRGB), double accumulation of them, the doubleRGB total and
    # central finite differences of the doubleRGB loss (the gradient oracle)
    fitted = "Aggregate(Lambertian(), CookTorrance())"
    truth = "Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))"
    loss = {}
    meta = {"fitted": fitted, "truth": truth, "grid": [[13, 8], [5, 6]], "metrics": {}}
    p0 = np.array([0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.1, 1.3])

    def ref_order(p):       # forward order -> the reference's per-lobe reversed run-time order (fact 14)
        return np.concatenate([p[:3][::-1], p[3:][::-1]])
    for name in METRICS:
        low = name in ("lowL2", "lowLog")
        so = (1, 6) if low else (5, 6)
        df = sph_desc((13, 8), so, end_out=(float(np.float32(2) * np.float32(np.pi)), float(np.float32(0.5) * np.float32(np.pi))) if low else None)
        dd = sph_desc((13, 8), so, real=np.float64, end_out=(2 * np.pi, 0.5 * np.pi) if low else None)
        N = 13 * 8 * so[0] * so[1]
        terms, tot = ref.loss(name, df, fitted, truth, 0, N)
        loss[name + "_terms"] = terms
        fd = []
        for j in range(len(p0)):
            h = 1e-6 * max(1.0, abs(p0[j]))
            pp, pm = p0.copy(), p0.copy()
            pp[j] += h; pm[j] -= h
            l = refd.loss_at(name, dd, fitted, truth, np.stack([ref_order(pp), ref_order(pm)]))
            fd.append((l[0] - l[1]) / (2 * h))
        meta["metrics"][name] = {"samples_out": list(so), "N": N, "float_total": float(tot), "double_sum_of_float_terms": float(terms.astype(np.float64).sum() / N),
                                 "double_total": float(refd.loss_at(name, dd, fitted, truth, ref_order(p0)[None])[0]), "fd_gradient": fd}
    # MERL-grid loss on a slice of bins
    terms, _ = ref.loss("nganL2", None, fitted, truth, 700000, 4096, want_total=False)
    loss["merl_nganL2_terms_700000"] = terms
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **loss)
    json.dump(meta, open(os.path.join(OUT, "losses.json"), "w"), indent=1)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
