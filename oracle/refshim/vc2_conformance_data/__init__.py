"""Stand-in for the reference's un-vendored `vc2_conformance_data` dependency (setup.py install_requires):
only the file-name constants that reference modules import at module level.  The data files themselves
(test pictures, the static filter analysis bundle) are NOT provided: nothing generated through this shim
may read them."""

POINTER_SPRITE_FILENAME = "/nonexistent/vc2_conformance_data/pointer.png"
NATURAL_PICTURES_FILENAMES = []
STATIC_FILTER_ANALYSIS_BUNDLE_FILENAME = "/nonexistent/vc2_conformance_data/static_filter_analysis.zip"
