cd $GRAFT_REPO_ROOT
python tools/trajan_profile.py 2>&1 | tail -40
