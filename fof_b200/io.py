"""Stage artefacts on either side of the path, in the reference's formats:

* ``export`` -- helper.py:224-237: the (K, 11) ``CleanedCluster.properties`` table as a FITS primary HDU.  astropy.io.fits
  is not required: a float64 image HDU is a 2880-byte card header plus big-endian data.
* ``dump_clusters`` / ``load_clusters`` -- the pickles pipeline.py:103-120,173-179 passes between stages.
  ``load_clusters`` also reads ``Cluster`` lists pickled by the reference itself (module ``cluster``) into this
  package's classes, so candidates.dat / cleaned_candidates.dat from an earlier reference run can be continued here.
* ``properties_table`` -- the same (K, 11) table straight from a device-resident ``pipeline.Catalogue`` without
  materialising K Python objects.
"""
import io as _io
import os
import pickle

import numpy as np

from . import cluster as _cluster

PROPERTY_COLUMNS = ("ra", "dec", "z", "cluster_mass", "vel_disp", "projected_radius", "total_absmag", "richness", "D",
                    "flag_poor", "gal_id")   # cluster.py:120-136


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "T" if value else "F"
    else:
        v = str(int(value))
    text = f"{key:<8}= {v:>20}"
    if comment:
        text += " / " + comment
    return text[:80].ljust(80)


def write_fits_array(path, arr, overwrite=False):
    """A 2-D float64 array as a FITS primary HDU (BITPIX = -64; NAXIS1 = columns, the fastest axis)."""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    if arr.ndim != 2:
        raise ValueError("expected a 2-D array")
    if os.path.exists(path) and not overwrite:
        raise OSError(f"File {path!r} already exists.")   # astropy's writeto refuses to overwrite too
    cards = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", -64, "array data type"),
             _card("NAXIS", 2, "number of array dimensions"), _card("NAXIS1", arr.shape[1]), _card("NAXIS2", arr.shape[0]),
             _card("EXTEND", True), "END".ljust(80)]
    header = "".join(cards).encode("ascii")
    header += b" " * (-len(header) % 2880)
    data = arr.astype(">f8").tobytes()
    data += b"\0" * (-len(data) % 2880)
    with open(path, "wb") as f:
        f.write(header + data)


def read_fits_array(path):
    """Inverse of write_fits_array (primary HDU, BITPIX -64, two axes)."""
    with open(path, "rb") as f:
        raw = f.read()
    keys, pos, done = {}, 0, False
    while not done:
        block = raw[pos:pos + 2880].decode("ascii")
        pos += 2880
        for k in range(0, 2880, 80):
            card = block[k:k + 80]
            if card.startswith("END"):
                done = True
                break
            if card[8:10] == "= ":
                keys[card[:8].strip()] = card[10:].split("/")[0].strip()
    if int(keys["BITPIX"]) != -64 or int(keys["NAXIS"]) != 2:
        raise ValueError("not a 2-D float64 image HDU")
    n1, n2 = int(keys["NAXIS1"]), int(keys["NAXIS2"])
    return np.frombuffer(raw, dtype=">f8", count=n1 * n2, offset=pos).reshape(n2, n1).astype(np.float64)


def export(clusters, name, path=""):
    """helper.py:224-237: write ``path + name + ".fits"`` and return the (K, N) property array."""
    n_prop = len(clusters[0].properties)
    export_arr = np.zeros((len(clusters), n_prop))
    for i, c in enumerate(clusters):
        export_arr[i] = c.properties
    write_fits_array(path + name + ".fits", export_arr)
    return export_arr


def properties_table(catalogue):
    """(K, 11) array in PROPERTY_COLUMNS order from a pipeline.Catalogue (no per-cluster Python objects)."""
    K = len(catalogue)
    out = np.zeros((K, len(PROPERTY_COLUMNS)))
    if K == 0:
        return out
    centre = catalogue.galaxy_arr[catalogue.centre_row]
    out[:, 0:3] = centre[:, 0:3]
    out[:, 3] = catalogue.props[:, 0]    # cluster_mass
    out[:, 4] = catalogue.props[:, 2]    # vel_disp
    out[:, 5] = catalogue.props[:, 4]    # projected_radius
    out[:, 6] = catalogue.props[:, 5]    # total_absmag
    out[:, 7] = np.diff(catalogue.offsets)
    out[:, 8] = catalogue.D
    out[:, 9] = 0.0                      # flag_poor is set by hand in the reference's workflow
    out[:, 10] = centre[:, 6]
    return out


def dump_clusters(clusters, file):
    """pipeline.py:103-104 / :119-120 / :178-179."""
    if hasattr(file, "write"):
        pickle.dump(clusters, file)
    else:
        with open(file, "wb") as f:
            pickle.dump(clusters, f)


class _ReferenceUnpickler(pickle.Unpickler):
    """Maps the reference's top-level module ``cluster`` (cluster.py) onto fof_b200.cluster."""

    def find_class(self, module, name):
        if module == "cluster" and name in ("Cluster", "CleanedCluster"):
            return getattr(_cluster, name)
        return super().find_class(module, name)


def load_clusters(file):
    """Read a cluster list pickled by this package or by the reference (pipeline.py:110-111,127-128)."""
    if hasattr(file, "read"):
        return _ReferenceUnpickler(file).load()
    with open(file, "rb") as f:
        return _ReferenceUnpickler(_io.BytesIO(f.read())).load()
