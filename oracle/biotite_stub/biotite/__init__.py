"""TEST INFRASTRUCTURE ONLY: minimal stand-in for the `biotite` package.

biotite (the reference's third-party dependency, pyproject.toml:38) is not
installable offline.  This stub implements only the ~12 symbols the reference
touches so that the UNMODIFIED reference under
/root/reference/src can be imported in this container to (a) validate the
oracle and (b) generate tests/golden/*.npz.  Nothing in springcraft_b200/
imports this package.
"""
__version__ = "0.0-stub"
