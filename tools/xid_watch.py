"""Print NVML Xid critical-error events for GPU 0 for a few seconds (debug aid)."""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
es = nv.nvmlEventSetCreate()
nv.nvmlDeviceRegisterEvents(h, nv.nvmlEventTypeXidCriticalError, es)
print("xid watcher armed", flush=True)
t_end = time.time() + float(sys.argv[1])
while time.time() < t_end:
    try:
        d = nv.nvmlEventSetWait(es, 1000)
        print("XID event: type=%d data(xid)=%d" % (d.eventType, d.eventData), flush=True)
    except nv.NVMLError as e:
        if e.value != nv.NVML_ERROR_TIMEOUT:
            print("nvml error", e, flush=True); break
