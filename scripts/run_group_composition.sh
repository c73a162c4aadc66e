#!/bin/bash
# mirrors i2vgen-xl/scripts/run_group_composition.sh of the reference
cd "$(dirname "$0")/.."
python -m mvoc_b200.composite \
    --template_config "${1:-configs/group_composite/template.yaml}" \
    --configs_json "${2:-configs/group_composite/group_config.json}"
