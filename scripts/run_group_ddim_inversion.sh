#!/bin/bash
# mirrors i2vgen-xl/scripts/run_group_ddim_inversion.sh of the reference
cd "$(dirname "$0")/.."
python -m mvoc_b200.inverse \
    --template_config "${1:-configs/group_inversion/template.yaml}" \
    --configs_json "${2:-configs/group_inversion/group_config.json}"
