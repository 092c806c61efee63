#!/bin/bash
# Builds the UNMODIFIED reference with cargo and pins this repo's oracle against it (see compare_with_real_sloth.py).
# usage: tools/compare_with_real_sloth.sh /path/to/rust-sloth        (needs cargo, gcc, python3 + numpy; no GPU)
set -euo pipefail
ref=${1:?path to a checkout of ecumene/rust-sloth}
here=$(cd "$(dirname "$0")/.." && pwd)
cargo build --release --manifest-path "$ref/Cargo.toml"
make -s -C "$here/oracle"
make -s -C "$here/rust-sloth_b200" libsloth_host.so
python3 "$here/tools/compare_with_real_sloth.py" "$ref/target/release/sloth" "$ref/models"
