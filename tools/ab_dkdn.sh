TAG=$1
MORE3D_DKDN_VARIANT=2 python tools/ab_dkdn.py /tmp/ref.npz > gpurun_out/${TAG}_ab2.json 2>gpurun_out/${TAG}_ab2.err
MORE3D_DKDN_VARIANT=0 python tools/ab_dkdn.py /tmp/new.npz /tmp/ref.npz > gpurun_out/${TAG}_ab0.json 2>gpurun_out/${TAG}_ab0.err
cat gpurun_out/${TAG}_ab2.json
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_ab0.json')); print({k:v for k,v in d.items() if k!='compare'}); print({k:(v['same_zero_nan_pattern'],v['max_rel_diff'],v['n_diff']) for k,v in d['compare'].items() if not (v['same_zero_nan_pattern'] and v['n_diff']==0)})"
tail -3 gpurun_out/${TAG}_ab0.err
