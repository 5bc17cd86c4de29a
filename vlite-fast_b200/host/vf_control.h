/*
 * Control-plane hooks of process_baseband: one-character commands on a UDP
 * multicast group (src/def.h:4-10; group 224.3.29.71, reader port 20000,
 * src/multicast.h:14-20), polled without blocking once per second and while
 * waiting for an observation (src/process_baseband.cu:764-768, 793-795,
 * 1081-1092; get_cmds / test_for_cmd, src/utils.c:174-220).
 */
#ifndef VF_CONTROL_H
#define VF_CONTROL_H
#ifdef __cplusplus
extern "C" {
#endif

#define VF_CMD_START 'S'
#define VF_CMD_STOP  'C'
#define VF_CMD_QUIT  'Q'
#define VF_CMD_EVENT 'E'
#define VF_CMD_NONE  'N'
#define VF_MC_GROUP "224.3.29.71"
#define VF_MC_READER_PORT 20000

/* UDP socket bound to port, joined to group when group is a multicast address (a unicast
 * address just binds), non-blocking.  Returns the descriptor or -1. */
int vf_mc_open (const char *group, int port);
int vf_mc_close (int sock);
int vf_mc_send (const char *group, int port, const char *msg, int len);
/* 1 if any queued datagram contains the command character (src/utils.c:174-186) */
int vf_test_for_cmd (int cmd, int sock);
/* cmds[0..4] = START, STOP, QUIT, EVENT, NONE seen in the next queued datagram (src/utils.c:188-220) */
void vf_get_cmds (int cmds[5], int sock);

#ifdef __cplusplus
}
#endif
#endif
