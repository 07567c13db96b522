#!/bin/bash
# like ab_lib.sh, with an environment assignment applied to every run:  bash tools/ab_lib_env.sh TAG VAR=val default v_x ...
TAG=${1:?tag}; ASSIGN=${2:?VAR=val}; shift 2
export "$ASSIGN"
exec bash "$(dirname "$0")/ab_lib.sh" "$TAG" "$@"
