"""Differential check of the reference-structure workflow against the LIVE reference (build container only; TEST
INFRASTRUCTURE).  python oracle/validate_workflow.py

site_analysis_b200.reference_workflow (distance matrices answered by the CPU oracle here -- there is no GPU in the build
container; tests/test_workflow.py runs the same logic on the GPU operator) against the unmodified
site_analysis.reference_workflow on a jittered, shifted rock-salt supercell: alignment translation / metrics / aligned
coordinates bit for bit for both metrics, a species subset, Nelder-Mead and (seeded) differential evolution, the error
messages of a failed optimisation and of an unknown algorithm, and the generated polyhedral sites (indices, labels,
reference centres, primed image shifts, unwrapped vertex coordinates) for the label / labels / use_reference_centers /
target_species variants."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, refcases
refcases.import_reference()
from pymatgen.core import Lattice, Structure
from site_analysis.reference_workflow import ReferenceBasedSites as RefRBS
from site_analysis.reference_workflow.structure_aligner import StructureAligner as RefAligner
import oracle
oracle.build()
from site_analysis_b200 import distances, SimpleStructure
from site_analysis_b200.reference_workflow import ReferenceBasedSites, StructureAligner
def cpu(c1,c2,l):
    c1=np.ascontiguousarray(c1,dtype=np.float64).reshape(-1,3); c2=np.ascontiguousarray(c2,dtype=np.float64).reshape(-1,3)
    if len(c1)==0 or len(c2)==0: return np.zeros((len(c1),len(c2)))
    return oracle.all_mic_distances(c1,c2,np.ascontiguousarray(l,dtype=np.float64))
distances.all_mic_distances=cpu
rng=np.random.default_rng(5)
ref=Structure.from_spacegroup(sg="Fm-3m", lattice=Lattice.cubic(5.6), species=["Na","Cl"], coords=[[0,0,0],[0.5,0.5,0.5]])*[2,2,2]
shift=np.array([0.03,-0.02,0.045])
tf=(np.array(ref.frac_coords)+shift+rng.normal(0,0.004,size=(len(ref),3)))%1.0
sp=[s.species_string for s in ref]
tgt=Structure(ref.lattice, sp, tf)
mref=SimpleStructure(ref.lattice.matrix, sp, ref.frac_coords); mtgt=SimpleStructure(ref.lattice.matrix, sp, tf)
ok=True
for kw in (dict(metric='rmsd'), dict(metric='max_dist'), dict(metric='rmsd', species=['Cl']),
           dict(metric='rmsd', algorithm='differential_evolution', minimizer_options={'seed': 3, 'maxiter': 30}), dict(metric='rmsd', algorithm='differential_evolution', minimizer_options={'seed': 3, 'popsize': 8}), dict(metric='rmsd', algorithm='simplex'),
           dict(metric='rmsd', tolerance=1e-6, minimizer_options={'maxiter': 400})):
    try:
        a=RefAligner().align(ref,tgt,**kw)
    except ValueError as e:
        try:
            StructureAligner().align(mref,mtgt,**kw); print(kw,'reference raised, mine did not'); ok=False
        except ValueError as e2:
            print(kw,'both raise:', str(e)==str(e2), str(e)); ok&=str(e)==str(e2)
        continue
    b=StructureAligner().align(mref,mtgt,**kw)
    same=np.array_equal(np.asarray(a[1]).view(np.int64), np.asarray(b[1]).view(np.int64)) and a[2]==b[2] and np.array_equal(np.array(a[0].frac_coords).view(np.int64), np.asarray(b[0].frac_coords).view(np.int64))
    print(kw, 'translation', a[1], 'identical' if same else 'DIFFERENT'); ok&=same
for kw in (dict(label=None, labels=[f"s{i}" for i in range(32)]), dict(label="x"), dict(use_reference_centers=False), dict(target_species=["Cl"])):
    r=RefRBS(ref,tgt,align=True,align_species=["Cl"]); m=ReferenceBasedSites(mref,mtgt,align=True,align_species=["Cl"])
    A=r.create_polyhedral_sites("Na","Cl",3.0,6,**kw); B=m.create_polyhedral_sites("Na","Cl",3.0,6,**kw)
    same=len(A)==len(B) and all(a.vertex_indices==b.vertex_indices and a.label==b.label and ((a.reference_center is None and b.reference_center is None) or np.array_equal(a.reference_center.view(np.int64), b.reference_center.view(np.int64))) and np.array_equal(a._pbc_image_shifts,b._pbc_image_shifts) and np.array_equal(a.vertex_coords.view(np.int64), b.vertex_coords.view(np.int64)) for a,b in zip(A,B))
    print(kw.keys(), len(A), 'identical' if same else 'DIFFERENT'); ok&=same
for bad in (dict(align_species=["Xx"]), ):
    for cls,(x,y) in ((RefRBS,(ref,tgt)),(ReferenceBasedSites,(mref,mtgt))):
        try: cls(x,y,align=True,**bad); print("no error")
        except ValueError as e: print(cls.__module__.split('.')[0], str(e))
print("ALL IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
