for p in 1 4 8; do echo "PARTS $p"; GCS_B200_PARTS=$p python scratch/e2e_probe.py; done
echo NOGRAPH; GCS_B200_NOGRAPH=1 GCS_B200_PARTS=4 python scratch/e2e_probe.py
