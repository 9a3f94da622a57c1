import json,sys
d=json.loads(open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read())
print('value %.4g feat/s  fi/s %.4g  ms/step %.3f  lk_ms %.3f  e2e %.4g' % (d['value'], d['feature_iterations_per_sec'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))
