"""Seeded synthetic COSMOS-like galaxy catalogues (BASELINE.json configs).

The reference ships no data (its sqlite DB is git-ignored, /root/reference/.gitignore:5), so every
test and benchmark runs on these mocks.  Output is the 7-column DataFrame that
/root/reference/FoF/pipeline.py:51-54 feeds to ``candidate_center_search``:
``ra, dec, redshift, z_lower, z_upper+, RMag, ID`` sorted by redshift.
"""
import numpy as np
import pandas as pd

C_KMS = 299792.458
COLUMNS = ["ra", "dec", "redshift", "z_lower", "z_upper+", "RMag", "ID"]


def _angular_diameter_distance(z, H0=70.0, Om0=0.3):
    """Flat LCDM D_A [Mpc] by 64-point Gauss-Legendre (mock geometry only, not a parity path)."""
    x, w = np.polynomial.legendre.leggauss(64)
    z = np.atleast_1d(np.asarray(z, dtype=float))
    zz = 0.5 * z[:, None] * (x[None, :] + 1.0)
    e = np.sqrt(Om0 * (1 + zz) ** 3 + (1 - Om0))
    dc = (C_KMS / H0) * 0.5 * z * np.sum(w[None, :] / e, axis=1)
    return dc / (1 + z)


def _sample_z(rng, n, zmin, zmax):
    """dN/dz ~ z^2 exp(-z/0.5), truncated to [zmin, zmax], by inverse CDF on a fine grid."""
    grid = np.linspace(zmin, zmax, 20001)
    pdf = grid ** 2 * np.exp(-grid / 0.5)
    cdf = np.cumsum(pdf)
    cdf = (cdf - cdf[0]) / (cdf[-1] - cdf[0])
    return np.interp(rng.random(n), cdf, grid)


def make_catalog(n, seed=20260101, kind="cosmos", density=50000.0, zmin=0.5, zmax=2.0,
                 cluster_fraction=0.2, richness_range=(20, 300), ra0=150.1, dec0=2.2,
                 n_blobs=None, as_frame=True, photoz_decimals=None):
    """Build one mock catalogue.

    kind = "cosmos"  : footprint sized for ``density`` galaxies/deg^2 around (ra0, dec0);
                        (1 - cluster_fraction) uniform background + Plummer clusters
                        (configs[0..2] of BASELINE.json, depending on n).
    kind = "wide"    : RA in [100, 150], Dec in [-20, 20] regardless of n (configs[3]; exercises the
                        GriSPy RA cell-box truncation).
    kind = "dense"   : all cluster members in ``n_blobs`` (<= 20) giant Plummer blobs plus a 20 %
                        background (configs[4], percolation stress).
    photoz_decimals  : round z_lower / z_upper to that many decimals, as published photo-z catalogues quote them.
                        Thousands of rows then share a column-4 value, which is what the reference's tracker keys on
                        (fof.py:238-246); None keeps the continuous (unique) values.
    """
    rng = np.random.default_rng(seed)
    n = int(n)
    if kind == "wide":
        ra_lo, ra_hi, dec_lo, dec_hi = 100.0, 150.0, -20.0, 20.0
    else:
        side = np.sqrt(n / density)
        half_ra = 0.5 * side / np.cos(np.deg2rad(dec0))
        ra_lo, ra_hi = ra0 - half_ra, ra0 + half_ra
        dec_lo, dec_hi = dec0 - 0.5 * side, dec0 + 0.5 * side

    if kind == "dense":
        nb = int(n_blobs or 12)
        n_cl_members = int(0.8 * n)
        rich = np.full(nb, n_cl_members // nb, dtype=np.int64)
        rich[: n_cl_members - rich.sum()] += 1
    else:
        target = int(cluster_fraction * n)
        mean_rich = 0.5 * (richness_range[0] + richness_range[1])
        n_cl = int(target / mean_rich * 1.3) + 8 if target > 0 else 0  # over-draw, then trim
        rich = rng.integers(richness_range[0], richness_range[1] + 1, size=n_cl)
        if n_cl:
            csum = np.cumsum(rich)
            while csum[-1] < target:
                rich = np.concatenate([rich, rng.integers(richness_range[0], richness_range[1] + 1, size=n_cl)])
                csum = np.cumsum(rich)
            keep = int(np.searchsorted(csum, target, side="left")) + 1
            rich = rich[:keep]
            rich[-1] -= int(rich.sum() - target)
            if rich[-1] <= 0:
                rich = rich[:-1]
    n_members = int(rich.sum()) if len(rich) else 0
    n_bg = n - n_members

    ra = np.empty(n)
    dec = np.empty(n)
    z = np.empty(n)
    ra[:n_bg] = rng.uniform(ra_lo, ra_hi, n_bg)
    # uniform on the sphere in Dec (matters only for the wide field)
    s_lo, s_hi = np.sin(np.deg2rad(dec_lo)), np.sin(np.deg2rad(dec_hi))
    dec[:n_bg] = np.rad2deg(np.arcsin(rng.uniform(s_lo, s_hi, n_bg)))
    z[:n_bg] = _sample_z(rng, n_bg, zmin, zmax)

    if n_members:
        n_cl = len(rich)
        mra = 0.05 * (ra_hi - ra_lo)
        mdec = 0.05 * (dec_hi - dec_lo)
        c_ra = rng.uniform(ra_lo + mra, ra_hi - mra, n_cl)
        c_dec = rng.uniform(dec_lo + mdec, dec_hi - mdec, n_cl)
        c_z = _sample_z(rng, n_cl, zmin + 0.02, zmax - 0.02)
        sigma_v = rng.uniform(300.0, 1000.0, n_cl)
        a_mpc = 0.2 / 0.7  # Plummer scale 0.2 Mpc/h proper
        theta_a = np.rad2deg(a_mpc / _angular_diameter_distance(c_z))  # deg
        owner = np.repeat(np.arange(n_cl), rich)
        uu = rng.uniform(0.0, 0.99, n_members)
        radius = theta_a[owner] * np.sqrt(uu / (1.0 - uu))
        phi = rng.uniform(0.0, 2 * np.pi, n_members)
        m_dec = c_dec[owner] + radius * np.cos(phi)
        m_ra = c_ra[owner] + radius * np.sin(phi) / np.cos(np.deg2rad(m_dec))
        vel = rng.normal(0.0, sigma_v[owner])
        m_z = c_z[owner] + vel * (1.0 + c_z[owner]) / C_KMS
        ra[n_bg:] = m_ra
        dec[n_bg:] = m_dec
        z[n_bg:] = np.clip(m_z, zmin, zmax + 0.02)

    z_lower = z - np.abs(rng.normal(0.0, 0.02 * (1 + z)))
    z_upper = z + np.abs(rng.normal(0.0, 0.02 * (1 + z)))
    if photoz_decimals is not None:  # after every draw, so the other columns do not depend on it
        z_lower = np.round(z_lower, int(photoz_decimals))
        z_upper = np.round(z_upper, int(photoz_decimals))
    rmag = rng.normal(-21.0, 1.0, n)
    gid = np.arange(n, dtype=np.float64)

    order = np.argsort(z, kind="stable")
    cols = [a[order] for a in (ra, dec, z, z_lower, z_upper, rmag, gid)]
    if not as_frame:
        return np.column_stack(cols)
    return pd.DataFrame(dict(zip(COLUMNS, cols)), columns=COLUMNS)


def make_catalog_torch(n, seed=20260104, kind="wide", density=50000.0, zmin=0.5, zmax=2.0, cluster_fraction=0.2,
                       richness_range=(20, 300), ra0=150.1, dec0=2.2, device="cuda"):
    """The same distributions as ``make_catalog`` drawn with torch's generator on a GPU, for tables too large to build
    on the host in a benchmark's set-up time (BASELINE.json configs[3]: 100M galaxies).  Returns an (n, 7) float64 CUDA
    tensor sorted by redshift.  Deterministic for a given seed and device type, so every rank of a multi-GPU run can
    build the identical table; it is NOT the numpy stream of ``make_catalog`` (parity fixtures use that one)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    f64 = torch.float64
    n = int(n)

    def uniform(lo, hi, m):
        return lo + (hi - lo) * torch.rand(m, dtype=f64, device=dev, generator=g)

    def normal(m):
        return torch.randn(m, dtype=f64, device=dev, generator=g)

    grid = torch.linspace(zmin, zmax, 20001, dtype=f64, device=dev)
    pdf = grid ** 2 * torch.exp(-grid / 0.5)
    cdf = torch.cumsum(pdf, 0)
    cdf = (cdf - cdf[0]) / (cdf[-1] - cdf[0])

    def sample_z(m, lo=None, hi=None):
        u = torch.rand(m, dtype=f64, device=dev, generator=g)
        if lo is not None:  # restrict to [lo, hi] through the CDF
            c_lo = cdf[torch.searchsorted(grid, torch.tensor(lo, dtype=f64, device=dev))]
            c_hi = cdf[torch.searchsorted(grid, torch.tensor(hi, dtype=f64, device=dev)).clamp(max=len(grid) - 1)]
            u = c_lo + (c_hi - c_lo) * u
        i = torch.searchsorted(cdf, u).clamp(1, len(grid) - 1)
        t = (u - cdf[i - 1]) / (cdf[i] - cdf[i - 1]).clamp(min=1e-300)
        return grid[i - 1] + t * (grid[i] - grid[i - 1])

    if kind == "wide":
        ra_lo, ra_hi, dec_lo, dec_hi = 100.0, 150.0, -20.0, 20.0
    else:
        side = float(np.sqrt(n / density))
        half_ra = 0.5 * side / np.cos(np.deg2rad(dec0))
        ra_lo, ra_hi, dec_lo, dec_hi = ra0 - half_ra, ra0 + half_ra, dec0 - 0.5 * side, dec0 + 0.5 * side
    target = int(cluster_fraction * n)
    mean_rich = 0.5 * (richness_range[0] + richness_range[1])
    n_cl = int(target / mean_rich)
    rich = torch.randint(richness_range[0], richness_range[1] + 1, (max(n_cl, 1),), device=dev, generator=g)
    n_members = int(rich.sum().item()) if n_cl else 0
    if n_members > n // 2:
        n_members, n_cl = 0, 0
    n_bg = n - n_members
    ra = torch.empty(n, dtype=f64, device=dev)
    dec = torch.empty(n, dtype=f64, device=dev)
    z = torch.empty(n, dtype=f64, device=dev)
    ra[:n_bg] = uniform(ra_lo, ra_hi, n_bg)
    s_lo, s_hi = float(np.sin(np.deg2rad(dec_lo))), float(np.sin(np.deg2rad(dec_hi)))
    dec[:n_bg] = torch.rad2deg(torch.asin(uniform(s_lo, s_hi, n_bg)))
    z[:n_bg] = sample_z(n_bg)
    if n_members:
        mra, mdec = 0.05 * (ra_hi - ra_lo), 0.05 * (dec_hi - dec_lo)
        c_ra = uniform(ra_lo + mra, ra_hi - mra, n_cl)
        c_dec = uniform(dec_lo + mdec, dec_hi - mdec, n_cl)
        c_z = sample_z(n_cl, zmin + 0.02, zmax - 0.02)
        sigma_v = uniform(300.0, 1000.0, n_cl)
        d_a = torch.from_numpy(_angular_diameter_distance(c_z.cpu().numpy())).to(dev)
        theta_a = torch.rad2deg((0.2 / 0.7) / d_a)
        owner = torch.repeat_interleave(torch.arange(n_cl, device=dev), rich)
        uu = uniform(0.0, 0.99, n_members)
        radius = theta_a[owner] * torch.sqrt(uu / (1.0 - uu))
        phi = uniform(0.0, 2 * np.pi, n_members)
        m_dec = c_dec[owner] + radius * torch.cos(phi)
        ra[n_bg:] = c_ra[owner] + radius * torch.sin(phi) / torch.cos(torch.deg2rad(m_dec))
        dec[n_bg:] = m_dec
        vel = normal(n_members) * sigma_v[owner]
        z[n_bg:] = (c_z[owner] + vel * (1.0 + c_z[owner]) / C_KMS).clamp(zmin, zmax + 0.02)
    z_lower = z - (normal(n) * 0.02 * (1 + z)).abs()
    z_upper = z + (normal(n) * 0.02 * (1 + z)).abs()
    rmag = -21.0 + normal(n)
    order = torch.argsort(z, stable=True)
    out = torch.empty((n, 7), dtype=f64, device=dev)
    for col, a in enumerate((ra, dec, z, z_lower, z_upper, rmag)):
        out[:, col] = a[order]
    out[:, 6] = order.to(f64)   # the generation index, as in make_catalog (ID = arange before the redshift sort)
    return out
