"""TEST INFRASTRUCTURE ONLY: describes a skirt_b200.configs parameter dict in the small description language of the
reference harness (oracle/ref_harness.cpp).  Imported by tests/ and bench.py's CPU legs only."""
from skirt_b200 import configs, simulation as sim


def reference_spec(p, threads=1, seed=4357, dustsamples=100, storeabs=None, packages=None, with_extra=False):
    """The same configuration in the small description language of the reference harness (oracle/ref_harness.cpp),
    plus the per-component luminosities and dust-mix tables that go with it.  Used only by tests/ and by
    bench.py's CPU legs."""
    lg = configs.wavelength_grid(p)
    mix = sim.InterstellarDustMix(lg)
    if storeabs is None:
        storeabs = 1 if p["sim"] == "pan" else 0
    lines = [f"sim {p['sim']}", f"threads {threads}", f"seed {seed}", f"packages {float(packages if packages is not None else p['packages'])!r}"]
    if p["sim"] == "oligo":
        lines.append("wavelengths " + " ".join(repr(float(v)) for v in p["wavelengths"]))
    else:
        a, b, k = p["loggrid"]; lines.append(f"loggrid {a!r} {b!r} {k}")
    g = p.get("grid"); extra = {}
    if g is None:
        n = p["n"]; gridline = f"grid cartesian {n} {n} {n} lin lin lin"
    elif g["kind"] in ("octtree", "bintree"):
        search = {"TopDown": 0, "Neighbor": 1, "Bookkeeping": 2}[g.get("searchMethod", "Neighbor")]
        gridline = f"grid {g['kind']} {g['minLevel']} {g['maxLevel']} {search} {g.get('maxMassFraction', 1e-6)!r} 0 {g.get('sampleCount', 100)}"
    elif g["kind"] == "voronoi":
        gridline = "grid voronoi file"; extra["particles"] = configs.sph_particles(g["particles"], p["box"], g.get("seed", 0x5eed0004))
    elif g["kind"] == "amesh":
        gridline = "grid amesh"; grid = configs.dust_grid(p, lg, mix)
        extra["amesh"] = configs.synthetic_amesh(p["box"], g["root"], g["depth"], g["frac"]); extra["ameshdust"] = grid.densityUnits
    else:
        raise ValueError(g["kind"])
    lines += ["box " + " ".join(repr(float(v)) for v in p["box"]), gridline, f"dustsamples {dustsamples}", f"storeabs {int(storeabs)}"]
    if p.get("selfabsorption"):
        lines.append("selfabs 1")

    def words(g):
        if g["geometry"] == "expdisk":
            w = f"expdisk {g['hR']!r} {g['hz']!r} {g.get('Rmax', 0.0)!r} {g.get('zmax', 0.0)!r}"
        else:
            w = f"sersic {g['index']!r} {g['Re']!r} {g.get('q', 1.0)!r}"
        sp = g.get("spiral")
        if sp:
            w += f" spiral {sp['arms']} {sp['pitch']!r} {sp['radius']!r} {sp['phase']!r} {sp['weight']!r} {sp['index']}"
        return w
    for s in p["stellar"]:
        lines.append("stellar " + words(s))
    for d in p["dust"]:
        if d["geometry"] == "mesh":
            lines.append(f"ameshdust {extra['ameshdust']!r}")
        else:
            lines.append(f"dust {d['tau']!r} {d['lam']!r} " + words(d))
    for i in p["instruments"]:
        w = f"instrument {i['kind']} {i['name']} {i['distance']!r} {i['inclination']!r} {i.get('azimuth', 0.0)!r} {i.get('positionAngle', 0.0)!r}"
        if i["kind"] != "sed":
            w += f" {i['Nxp']} {i['fovxp']!r} {i['Nyp']} {i['fovyp']!r}"
        lines.append(w)
    out = ("\n".join(lines) + "\n", configs.luminosities(p, lg), [(mix.kappaabs, mix.kappasca, mix.asymmpar) for _ in p["dust"]])
    return out + (extra,) if with_extra else out
