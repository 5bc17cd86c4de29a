#include <arpa/inet.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>
#include "vf_control.h"

int vf_mc_open (const char *group, int port)
{
  int sock = socket (AF_INET, SOCK_DGRAM, 0);
  if (sock < 0) return -1;
  int one = 1;
  setsockopt (sock, SOL_SOCKET, SO_REUSEADDR, &one, sizeof (one));
  struct sockaddr_in addr;
  memset (&addr, 0, sizeof (addr));
  addr.sin_family = AF_INET;
  addr.sin_port = htons ((unsigned short) port);
  addr.sin_addr.s_addr = htonl (INADDR_ANY);
  if (bind (sock, (struct sockaddr *) &addr, sizeof (addr)) < 0) { close (sock); return -1; }
  struct in_addr g;
  if (group && inet_aton (group, &g) && IN_MULTICAST (ntohl (g.s_addr))) {
    struct ip_mreq mreq;
    mreq.imr_multiaddr = g;
    mreq.imr_interface.s_addr = htonl (INADDR_ANY);
    if (setsockopt (sock, IPPROTO_IP, IP_ADD_MEMBERSHIP, &mreq, sizeof (mreq)) < 0) { close (sock); return -1; }
  }
  fcntl (sock, F_SETFL, fcntl (sock, F_GETFL, 0) | O_NONBLOCK);
  return sock;
}

int vf_mc_close (int sock) { return sock >= 0 ? close (sock) : 0; }

int vf_mc_send (const char *group, int port, const char *msg, int len)
{
  int sock = socket (AF_INET, SOCK_DGRAM, 0);
  if (sock < 0) return -1;
  struct sockaddr_in addr;
  memset (&addr, 0, sizeof (addr));
  addr.sin_family = AF_INET;
  addr.sin_port = htons ((unsigned short) port);
  if (!inet_aton (group, &addr.sin_addr)) { close (sock); return -1; }
  unsigned char ttl = 3, loop = 1;
  setsockopt (sock, IPPROTO_IP, IP_MULTICAST_TTL, &ttl, sizeof (ttl));
  setsockopt (sock, IPPROTO_IP, IP_MULTICAST_LOOP, &loop, sizeof (loop));
  int n = (int) sendto (sock, msg, (size_t) len, 0, (struct sockaddr *) &addr, sizeof (addr));
  close (sock);
  return n;
}

int vf_test_for_cmd (int cmd, int sock)
{
  char buf[32];
  int found = 0;
  if (sock < 0) return 0;
  for (;;) {
    int n = (int) read (sock, buf, sizeof (buf));
    if (n <= 0) break;
    for (int i = 0; i < n; ++i) if (buf[i] == cmd) found = 1;
  }
  return found;
}

void vf_get_cmds (int cmds[5], int sock)
{
  char buf[32];
  for (int i = 0; i < 5; ++i) cmds[i] = 0;
  if (sock < 0) return;
  int n = (int) read (sock, buf, sizeof (buf));
  for (int i = 0; i < n; ++i) {
    if (buf[i] == VF_CMD_START) cmds[0] = 1;
    else if (buf[i] == VF_CMD_STOP) cmds[1] = 1;
    else if (buf[i] == VF_CMD_QUIT) cmds[2] = 1;
    else if (buf[i] == VF_CMD_EVENT) cmds[3] = 1;
    else if (buf[i] == VF_CMD_NONE) cmds[4] = 1;
  }
}
